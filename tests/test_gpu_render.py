"""GPU parity of the fused operator render(rays_o, rays_d, near, far, chunk) -> {rgb, depth, acc, weights}
against the golden chain vectors (produced by the real reference) and the oracle; size-independent
properties at the full 1008x756 frame; host-buffer entry; chunk / shard invariance; the reference-signature shims."""
import numpy as np
import pytest
import torch

import render_oracle as O
import tgtc_style_b200 as T
from helpers import FERN_FULL, golden, knife_edge_mask, small_rays, weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["w1", "w0"])
def test_render_fp32_end_to_end_vs_reference_vectors(renderer_fp32, kind):
    """fp32 mode: strict max <= 1e-3 on rgb/depth/acc, END TO END (not teacher-forced), vs the reference."""
    g = golden("chain_" + kind)
    wc, wf = weights(kind)
    renderer_fp32.set_weights(wc, wf)
    out = renderer_fp32.render(g["rays_o"], g["rays_d"], 0., 1., n_samples=64, n_fine=64, extras=True)
    tol = 1e-3
    assert np.abs(out["rgb"].cpu().numpy() - g["rgb"]).max() <= tol
    assert np.abs(out["depth"].cpu().numpy() - g["depth"]).max() <= tol
    assert np.abs(out["acc"].cpu().numpy() - g["acc"]).max() <= tol
    assert np.abs(out["rgb_coarse"].cpu().numpy() - g["rgb_coarse"]).max() <= tol
    assert np.abs(out["weights"].cpu().numpy() - g["weights"]).max() <= tol
    assert np.abs(out["ts_fine"].cpu().numpy() - g["ts_fine"]).max() <= 5e-3   # intermediate, not in the contract


def test_render_bf16_end_to_end_default_init(renderer_bf16):
    """bf16 mode on the literal BASELINE weight set (W0: seed-0 default init): END-TO-END max <= 1e-2 on
    rgb/depth/acc against the reference's own vectors."""
    g = golden("chain_w0")
    wc, wf = weights("w0")
    renderer_bf16.set_weights(wc, wf)
    out = renderer_bf16.render(g["rays_o"], g["rays_d"], 0., 1., n_samples=64, n_fine=64)
    tol = 1e-2
    assert np.abs(out["rgb"].cpu().numpy() - g["rgb"]).max() <= tol
    assert np.abs(out["depth"].cpu().numpy() - g["depth"]).max() <= tol
    assert np.abs(out["acc"].cpu().numpy() - g["acc"]).max() <= tol


@pytest.mark.parametrize("mode", ["f16", "bf16"])
def test_render_tc_teacher_forced_protocol(mode):
    """Tensor-core modes on the sigma-recalibrated set W1 (sigma ~ N(0,30^2): deliberately hard, every sigma error is
    amplified by the compositing), SURVEY H1 protocol: fine pass teacher-forced on the reference's ts_fine, knife-edge rays
    (oracle moves > 1e-2 under a small sigma perturbation, helpers.knife_edge_mask) flagged, < 0.5 % of the rays.
    fp16 operands: max <= 1e-2 over the non-flagged rays (north star).  bf16 operands: the format itself leaves ~0.3 % of such
    rays above 1e-2 (tools/precision_study.py, tests/test_gpu_parity_fullsize.py), so on these 160 rays: >= 98 % within
    1e-2 (at most three rays above), raw max <= 3e-2."""
    r = T.NerfRenderer(device="cuda:0", mode=mode)
    g = golden("chain_w1")
    wc, wf = weights("w1")
    r.set_weights(wc, wf)
    rs = r.nerf_forward_rays(T.NET_FINE, g["rays_o"], g["rays_d"], g["ts_fine"], 128, 0., 1.)
    rgb, depth, w, acc = r.composite(t_values=g["ts_fine"], rgbsigma=rs)
    flagged = knife_edge_mask(torch.from_numpy(g["sigma_fine"]), torch.from_numpy(g["ts_fine"])).numpy()
    ok = ~flagged
    assert flagged.mean() <= 0.005
    e_rgb = np.abs(rgb.cpu().numpy() - g["rgb"]).max(-1)
    e_acc = np.abs(acc.cpu().numpy() - g["acc"])
    e_dep = np.abs(depth.cpu().numpy() - g["depth"])
    print("%s teacher-forced: max drgb %.2e dacc %.2e ddepth %.2e, frac>1e-2 %.4f" % (mode, e_rgb[ok].max(), e_acc[ok].max(), e_dep[ok].max(), (e_rgb > 1e-2).mean()))
    for e in (e_rgb, e_acc, e_dep):
        if mode == "f16":
            assert e[ok].max() <= 1e-2
            assert e.mean() <= 3e-4
        else:
            assert (e[ok] > 1e-2).sum() <= 3
            assert e[ok].max() <= 3e-2
            assert e.mean() <= 2e-3
    out = r.render(g["rays_o"], g["rays_d"], 0., 1., n_samples=64, n_fine=64)
    assert np.abs(out["rgb"].cpu().numpy() - g["rgb"]).mean() <= (1e-3 if mode == "f16" else 5e-3)      # end to end (resampling moves ts_fine)
    r.close()


def test_render_host_chunk_and_shard_invariance(renderer_fp32):
    """rays are independent: chunked passes, host-buffer entry and 2/3/8-way ray shards all give the same BITS."""
    wc, wf = weights("w1")
    renderer_fp32.set_weights(wc, wf)
    ro, rd = small_rays()
    sel = np.arange(0, ro.shape[0], 331)[:500]
    o, d = torch.from_numpy(ro[sel]), torch.from_numpy(rd[sel])
    full = renderer_fp32.render(o, d, 0., 1., extras=True)
    chunked = renderer_fp32.render(o, d, 0., 1., chunk=128, extras=True)
    host = renderer_fp32.render_host(o.pin_memory(), d.pin_memory(), 0., 1., extras=True, want_weights=True)
    for k in full:
        assert torch.equal(full[k], chunked[k]), k
        assert torch.equal(full[k].cpu(), host[k]), k
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            b, e = T.shard_range(500, r, world)
            parts.append(renderer_fp32.render(o[b:e], d[b:e], 0., 1.))
        for k in ("rgb", "depth", "acc", "weights"):
            assert torch.equal(torch.cat([p[k] for p in parts], 0), full[k]), (world, k)


def test_render_frame_properties_full_size(renderer_bf16):
    """1008x756 (BASELINE config 2) through render_frame (ray-gen fused): properties the domain offers."""
    H, W, f = FERN_FULL
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
    n = 1008 * 96                                          # 96 image rows (the whole frame is timed in bench.py)
    out = renderer_bf16.render_frame(H, W, K, np.eye(4)[:3, :4], pix_begin=1008 * 300, n=n, extras=True, want_weights=True)
    ts = out["ts_fine"]
    assert (ts[:, 1:] >= ts[:, :-1]).all() and (ts[:, 0] == 0).all() and (ts[:, -1] == 1).all()
    assert torch.isfinite(out["rgb"]).all() and (out["rgb"] >= 0).all() and (out["rgb"] <= 1 + 1e-5).all()
    assert (out["weights"] >= 0).all() and (out["acc"] <= 1 + 1e-4).all()
    assert torch.allclose(out["weights"].sum(-1), out["acc"], atol=1e-5)
    assert (out["depth"] >= -1e-6).all() and (out["depth"] <= 1 + 1e-4).all()
    # same pixels through explicit rays: identical bits
    ro, rd = renderer_bf16.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=1008 * 300, n=n)
    out2 = renderer_bf16.render(ro, rd, 0., 1.)
    assert torch.equal(out2["rgb"], out["rgb"]) and torch.equal(out2["depth"], out["depth"])


def test_empty_and_tiny_inputs(renderer_fp32):
    wc, wf = weights("w1")
    renderer_fp32.set_weights(wc, wf)
    out = renderer_fp32.render(torch.zeros(0, 3), torch.zeros(0, 3), 0., 1.)
    assert out["rgb"].shape == (0, 3)
    ro, rd = small_rays()
    one = renderer_fp32.render(ro[:1], rd[:1], 0., 1.)
    three = renderer_fp32.render(ro[:3], rd[:3], 0., 1.)
    assert torch.equal(one["rgb"][0], three["rgb"][0])
    coarse_only = renderer_fp32.render(ro[:3], rd[:3], 0., 1., n_fine=0, extras=True)
    assert torch.equal(coarse_only["rgb"], coarse_only["rgb_coarse"])


def test_unset_weights_fail_loudly():
    r = T.NerfRenderer(device="cuda:0", mode="fp32")
    with pytest.raises(T.TgtcError, match="weights"):
        r.render(torch.zeros(4, 3), torch.ones(4, 3), 0., 1.)
    r.close()


def test_reference_signature_shims(renderer_fp32):
    """the four injected callables driven exactly as rendering.py:27-51 drives them."""
    g = golden("chain_w1")
    wc, wf = weights("w1")
    renderer_fp32.set_weights(wc, wf)
    fns = T.make_callables(renderer_fp32)
    dev = renderer_fp32.device
    rays_o, rays_d = torch.from_numpy(g["rays_o"]).to(dev), torch.from_numpy(g["rays_d"]).to(dev)
    n = rays_o.shape[0]
    pts, ts = fns["sampling_pts_uniform"](rays_o=rays_o, rays_d=rays_d, N_samples=64, near=0., far=1.)
    ret = fns["model_forward"](pts=pts, dirs=rays_d.unsqueeze(1).expand([n, 64, 3]))
    assert set(ret) == {"rgb", "base_remap", "pts", "sigma", "dirs"}
    rgb_exp, t_exp, weights_c = fns["alpha_composition"](ret["rgb"], ret["sigma"], ts, 0)
    pts_fine, ts_fine = fns["sampling_pts_fine_torch"](rays_o, rays_d, ts, weights_c, 64)
    ret_f = fns["model_forward_fine"](pts=pts_fine, dirs=rays_d.unsqueeze(1).expand([n, 128, 3]))
    rgb_f, t_f, _ = fns["alpha_composition"](ret_f["rgb"], ret_f["sigma"], ts_fine, 0)
    assert np.abs(rgb_f.cpu().numpy() - g["rgb"]).max() <= 1e-3
    assert np.abs(t_f.cpu().numpy() - g["depth"]).max() <= 1e-3
    assert ret_f["base_remap"].shape == (n, 128, 256) and ret_f["pts"].shape == (n, 128, 63) and ret_f["dirs"].shape == (n, 128, 27)
    # sigma_noise_std > 0 draws from torch's generator like utils.py:372-374 and still composites
    torch.manual_seed(0)
    noisy = fns["alpha_composition"](ret["rgb"], ret["sigma"], ts, 1.0)
    assert torch.isfinite(noisy[0]).all()


def test_render_path_frames_match_single_frame_render(renderer_bf16):
    """render_path_sharded (world 1) == render_frame per pose, rays generated on device, frames kept on device."""
    import tgtc_style_b200 as T
    H, W, f = 48, 64, 52.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    poses = [np.eye(4)[:3, :4], np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.05], [0, 0, 1, 0.02]], dtype=np.float64)]
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    frames = dict(T.render_path_sharded(renderer_bf16, H, W, K, poses, split="rows"))
    assert sorted(frames) == [0, 1]
    for i, pose in enumerate(poses):
        ref = renderer_bf16.render_frame(H, W, K, pose)
        for k in ("rgb", "depth", "acc"):
            assert torch.equal(frames[i][k], ref[k])
    assert not torch.equal(frames[0]["rgb"], frames[1]["rgb"])


def test_cal_geometry_outputs(renderer_fp32, tmp_path):
    """rendering.py:54,:66-81: coor_map = t*d + o and the per-frame npz the 2-D temporal trainer reads."""
    import tgtc_style_b200 as T
    H, W, f = 24, 32, 26.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    poses = [np.eye(4)[:3, :4], np.array([[1, 0, 0, 0.1], [0, 1, 0, 0.0], [0, 0, 1, 0.05]], dtype=np.float64)]
    wc, wf = weights("w1")
    renderer_fp32.set_weights(wc, wf)
    rgb_map, t_map = T.cal_geometry(renderer_fp32, H, W, K, poses, hwf=[H, W, f], sv_path=str(tmp_path))
    assert rgb_map.shape == (2, H, W, 3) and t_map.shape == (2, H, W, 1)
    ro, rd = O.make_rays(H, W, f, poses[1])
    ref = O.render_chain(wc, wf, ro, rd, 0., 1., 64, 64, 1024)
    np.testing.assert_allclose(t_map[1].reshape(-1), ref["depth"].numpy(), atol=1e-3)
    g = np.load(str(tmp_path / "geometry_00001.npz"))
    assert set(g.files) == {"coor_map", "cps", "hwf", "near", "far"} and g["coor_map"].shape == (H, W, 3)
    coor_ref = ref["depth"].numpy()[:, None] * rd + ro
    np.testing.assert_allclose(g["coor_map"].reshape(-1, 3), coor_ref, atol=2e-3)
    full = np.load(str(tmp_path / "geometry.npz"))
    assert full["coor_map"].shape == (2, H, W, 3) and (tmp_path / "rgb_00000.png").exists() and (tmp_path / "depth_00001.png").exists()


@pytest.mark.parametrize("n", [1, 2, 3, 127, 129, 1000, 4097])
def test_render_bf16_ragged_sizes_match_prefix_of_larger_batch(renderer_bf16, n):
    """tile quads with padding tiles / half-filled last tiles (CTA pairs work on 4 tiles at a time): rendering the first n
    rays alone gives bit-identical results to rendering them inside a larger batch."""
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    ro, rd = small_rays()
    sel = np.linspace(0, ro.shape[0] - 1, 5000).astype(np.int64)
    big = renderer_bf16.render(ro[sel], rd[sel], 0., 1., extras=True)
    small = renderer_bf16.render(ro[sel[:n]], rd[sel[:n]], 0., 1., extras=True)
    for k in ("rgb", "depth", "acc", "weights", "rgb_coarse", "ts_fine"):
        assert torch.equal(small[k], big[k][:n]), k
    assert torch.isfinite(small["rgb"]).all()


@pytest.mark.parametrize("mode", ["f16", "bf16"])
@pytest.mark.parametrize("n", [1, 2, 3, 129, 1001])
def test_fused_compositing_is_bit_identical_to_the_standalone_kernel(mode, n):
    """K5 fused into the tcgen05 MLP kernel (a 128-sample tile = one fine ray / two coarse rays) and K6+K7 fused into the fine
    kernel's producer warps, against the same render with the stand-alone composite / sample_fine kernels: every output (incl.
    ts_fine), bit for bit -- odd ray counts (half-filled last coarse tile),
    white background, weights requested."""
    from tgtc_style_b200 import _lib
    r = T.NerfRenderer(device="cuda:0", mode=mode)
    wc, wf = weights("w1")
    r.set_weights(wc, wf)
    ro, rd = small_rays()
    sel = np.linspace(0, ro.shape[0] - 1, n).astype(np.int64)
    lib = _lib.load()
    outs = []
    for off in (0, 1):
        lib.tgtc_debug_no_fused_composite(off)
        lib.tgtc_debug_no_fused_sample_fine(off)     # likewise the resampling: fused into the fine kernel's producers vs its own kernel
        try:
            a = r.render(ro[sel], rd[sel], 0., 1., extras=True, want_weights=True)
            b = r.render(ro[sel], rd[sel], 0., 1., extras=True, want_weights=True, white_bkgd=True)
            torch.cuda.synchronize()
            outs.append(({k: v.clone() for k, v in a.items()}, {k: v.clone() for k, v in b.items()}))
        finally:
            lib.tgtc_debug_no_fused_composite(0)
            lib.tgtc_debug_no_fused_sample_fine(0)
    for (fa, fb), (sa, sb) in [(outs[0], outs[1])]:
        for k in fa:
            assert torch.equal(fa[k], sa[k]), ("plain", k)
            assert torch.equal(fb[k], sb[k]), ("white_bkgd", k)
    assert torch.isfinite(outs[0][0]["rgb"]).all()
    # the fused path needs ~4 B/sample of workspace instead of 20
    assert lib.tgtc_render_workspace_bytes_mode(r.mode, 4096, 64, 64, 0) * 4 < lib.tgtc_render_workspace_bytes(4096, 64, 64, 0)
    r.close()
