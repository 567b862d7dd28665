"""Per-ray style head on the GPU: tcgen05 chain kernels against the oracle restatement of render_style's loop
body (rendering.py:118-178) and the golden vectors produced by the imported reference."""
import numpy as np
import pytest
import torch

import render_oracle as O
from helpers import golden, small_rays

pytestmark = pytest.mark.gpu


def _setup(r):
    g = golden("style_chain")
    ro, rd = small_rays()
    w0c, w0f = O.init_linear_like_reference(0)
    probe = g["probe_index"]
    wc, wf = O.recalibrate_sigma(w0c, ro[probe], rd[probe]), O.recalibrate_sigma(w0f, ro[probe], rd[probe])
    cs, ws = O.init_style_like_reference(1)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    return g, ro, rd, (wc, wf, cs, ws)


def test_style_render_vs_reference_golden(renderer_bf16):
    r = renderer_bf16
    g, ro, rd, _ = _setup(r)
    sel = g["ray_index"]
    out = r.render_style(ro[sel], rd[sel], g["latents"], extras=True, want_weights=True)
    torch.cuda.synchronize()
    for k in ("rgb", "rgb_coarse", "depth", "weights_coarse"):
        assert torch.isfinite(out[k]).all(), k
    # coarse pass: sample positions are deterministic, so this is a teacher-forced comparison of
    # trunk -> module 1 -> module 2 -> compositing in bf16 against the fp32 reference
    ec = (out["rgb_coarse"].cpu().numpy() - g["rgb_coarse"])
    print("style coarse rgb: max %.3e mean %.3e" % (np.abs(ec).max(), np.abs(ec).mean()))
    assert np.abs(ec).mean() <= 3e-3
    assert (np.abs(ec).max(-1) <= 1e-2).mean() >= 0.97       # knife-edge rays (delta_last = 1e10) excepted
    ef = (out["rgb"].cpu().numpy() - g["rgb"])
    print("style fine rgb: max %.3e mean %.3e" % (np.abs(ef).max(), np.abs(ef).mean()))
    assert np.abs(ef).mean() <= 1e-2


def test_style_render_vs_oracle_other_latent_and_chunks(renderer_bf16):
    r = renderer_bf16
    g, ro, rd, (wc, wf, cs, ws) = _setup(r)
    sel = np.linspace(0, ro.shape[0] - 1, 300).astype(np.int64)
    lat = torch.randn(32, generator=torch.Generator().manual_seed(9)) * 0.7
    ref = O.render_style_chain(wc, wf, cs, ws, ro[sel], rd[sel], lat.unsqueeze(0).expand(len(sel), 32))
    a = r.render_style(ro[sel], rd[sel], lat, extras=True)
    b = r.render_style(ro[sel], rd[sel], lat, chunk=128, extras=True)        # passes of 128 rays: same result, bit for bit
    torch.cuda.synchronize()
    for k in ("rgb", "rgb_coarse", "depth", "acc"):
        assert torch.equal(a[k], b[k]), k
    ec = (a["rgb_coarse"].cpu() - ref["rgb_coarse"]).abs()
    assert ec.mean().item() <= 3e-3 and (ec.max(-1)[0] <= 1e-2).float().mean().item() >= 0.97
    # the latent matters: a different latent gives a different image
    c = r.render_style(ro[sel], rd[sel], -lat)
    assert (c["rgb"] - a["rgb"]).abs().mean().item() > 1e-4


def test_style_requires_weights_and_single_latent(renderer_bf16):
    import tgtc_style_b200 as T
    r2 = T.NerfRenderer(device="cuda:0", mode="bf16")
    ro, rd = small_rays()
    w0c, w0f = O.init_linear_like_reference(0)
    r2.set_weights(w0c, w0f)
    with pytest.raises(T.TgtcError):
        r2.render_style(ro[:64], rd[:64], torch.zeros(32))
    r2.close()


def test_style_per_ray_latents_in_one_call(renderer_bf16):
    """[N,32] latents that change along the batch go through ONE library call (tgtc_render_style_rays: per-ray effective biases):
    == the runs rendered separately with their own latent (different bias kernels: fp32 summation order only), and against the
    oracle with fully per-ray latents (rendering.py:125-127)."""
    r = renderer_bf16
    g, ro, rd, (wc, wf, cs, ws) = _setup(r)
    sel = np.linspace(0, ro.shape[0] - 1, 96).astype(np.int64)
    gen = torch.Generator().manual_seed(2)
    la, lb = torch.randn(32, generator=gen), torch.randn(32, generator=gen)
    lat = torch.cat([la.expand(40, 32), lb.expand(56, 32)], 0)
    l0 = r.launch_count()
    both = r.render_style(ro[sel], rd[sel], lat, extras=True)
    n_launch = r.launch_count() - l0
    a = r.render_style(ro[sel[:40]], rd[sel[:40]], la, extras=True)
    b = r.render_style(ro[sel[40:]], rd[sel[40:]], lb, extras=True)
    torch.cuda.synchronize()
    assert n_launch <= 14                       # one pass: bias kernel + 2 x (trunk, module 1, module 2, compositing) + sampling
    for k in ("rgb_coarse", "depth_coarse"):     # the coarse pass is teacher-forced by construction
        assert torch.allclose(both[k][:40], a[k], atol=2e-5) and torch.allclose(both[k][40:], b[k], atol=2e-5), k
    assert (both["rgb"][:40] - a["rgb"]).abs().mean().item() <= 1e-3
    # every ray its own latent
    lat_r = torch.randn(96, 32, generator=gen) * 0.7
    ref = O.render_style_chain(wc, wf, cs, ws, ro[sel], rd[sel], lat_r)
    out = r.render_style(ro[sel], rd[sel], lat_r, extras=True)
    torch.cuda.synchronize()
    ec = (out["rgb_coarse"].cpu() - ref["rgb_coarse"]).abs()
    print("per-ray latents coarse rgb: max %.3e mean %.3e" % (ec.max().item(), ec.mean().item()))
    assert ec.mean().item() <= 3e-3 and (ec.max(-1)[0] <= 1e-2).float().mean().item() >= 0.97


@pytest.mark.parametrize("per_ray", [False, True])
def test_style_render_f16_coarse_pass_max_bound(per_ray):
    """fp16-operand mode of the stylised render (trunk, module 1, module 2 and the feature tiles between them): the coarse pass
    -- teacher-forced by construction -- within 1e-2 on every ray that is not knife-edge, against the oracle."""
    import tgtc_style_b200 as T
    from helpers import knife_edge_mask
    r = T.NerfRenderer(device="cuda:0", mode="f16")
    g, ro, rd, (wc, wf, cs, ws) = _setup(r)
    sel = np.linspace(0, ro.shape[0] - 1, 600).astype(np.int64)
    gen = torch.Generator().manual_seed(9)
    lat = (torch.randn(len(sel), 32, generator=gen) if per_ray else torch.randn(32, generator=gen).unsqueeze(0).expand(len(sel), 32)) * 0.7
    ref = O.render_style_chain(wc, wf, cs, ws, ro[sel], rd[sel], lat)
    out = r.render_style(ro[sel], rd[sel], lat if per_ray else lat[0], extras=True)
    torch.cuda.synchronize()
    pts, ts = O.sample_uniform(torch.from_numpy(ro[sel]), torch.from_numpy(rd[sel]), 64, 0., 1.)
    sig = O.nerf_forward(wc, pts, torch.from_numpy(rd[sel]).unsqueeze(1).expand(len(sel), 64, 3))["sigma"]
    flagged = knife_edge_mask(sig, ts.contiguous())
    e = torch.stack([(out["rgb_coarse"].cpu() - ref["rgb_coarse"]).abs().max(-1)[0], (out["depth_coarse"].cpu() - ref["depth_coarse"]).abs()], 0).max(0)[0]
    print("style f16 coarse (per_ray=%s): flagged %.3f%% max(not flagged) %.2e mean %.2e" % (per_ray, 100 * flagged.float().mean(), e[~flagged].max(), e.mean()))
    assert flagged.float().mean().item() < 0.005
    assert e[~flagged].max().item() <= 1e-2
    ef = (out["rgb"].cpu() - ref["rgb"]).abs()
    assert ef.mean().item() <= 1e-3              # fine pass: end to end (resampled on its own coarse weights)
    r.close()
