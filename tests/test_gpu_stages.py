"""GPU parity, stage by stage (teacher-forced on the oracle's / golden inputs), through the C ABI:
K1 ray-gen+NDC, K2 uniform sampling, K5 compositing, K6+K7 inverse-CDF resampling + sort."""
import numpy as np
import pytest
import torch

import render_oracle as O
from helpers import FERN_FULL, FERN_SMALL, golden, small_rays

pytestmark = pytest.mark.gpu


def K_of(H, W, f):
    return np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]], np.float64)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("tag,hwf", [("small_identity", FERN_SMALL), ("full_identity", FERN_FULL), ("full_spiral17", FERN_FULL)])
def test_raygen_bit_exact(renderer_fp32, tag, hwf):
    g = golden("rays")
    H, W, f = hwf
    c2w = g[tag + "_c2w"]
    ro, rd = renderer_fp32.raygen(H, W, K_of(H, W, f), c2w)
    ro, rd = ro.cpu().numpy(), rd.cpu().numpy()
    idx = g[tag + "_idx"]
    assert np.array_equal(ro[idx], g[tag + "_o"]) and np.array_equal(rd[idx], g[tag + "_d"])   # reference vectors
    oo, od = O.make_rays(H, W, f, c2w)                                                          # oracle, every pixel
    assert np.array_equal(ro, oo) and np.array_equal(rd, od)


def test_raygen_ranges_and_no_ndc(renderer_fp32):
    H, W, f = FERN_SMALL
    g = golden("rays")
    c2w = g["spiral_poses"][41]
    oo, od = O.make_rays(H, W, f, c2w)
    # ragged pixel ranges (not multiples of 4, unaligned starts) tile the frame exactly
    cuts = [0, 1, 7, 1000, 1003, 50001, H * W - 5, H * W]
    for b, e in zip(cuts, cuts[1:]):
        ro, rd = renderer_fp32.raygen(H, W, K_of(H, W, f), c2w, pix_begin=b, n=e - b)
        assert np.array_equal(ro.cpu().numpy(), oo[b:e]) and np.array_equal(rd.cpu().numpy(), od[b:e])
    ro, rd = renderer_fp32.raygen(H, W, K_of(H, W, f), c2w, ndc=False, pixel_alignment=True)
    po, pd = O.make_rays(H, W, f, c2w, ndc=False, pixel_alignment=True)
    assert np.array_equal(ro.cpu().numpy(), po) and np.array_equal(rd.cpu().numpy(), pd)
    ro, rd = renderer_fp32.raygen(H, W, K_of(H, W, f), c2w, pix_begin=5, n=0)
    assert ro.shape == (0, 3)


# ------------------------------------------------------------------ K2
@pytest.mark.parametrize("S,near,far", [(64, 0., 1.), (64, 0., 1.05), (128, 2., 6.), (33, 0.3, 0.9)])
def test_sample_uniform_bit_exact(renderer_fp32, S, near, far):
    ro, rd = small_rays()
    o, d = ro[::977][:150], rd[::977][:150]
    pts, ts = renderer_fp32.sample_uniform(o, d, S, near, far)
    p_ref, t_ref = O.sample_uniform(o, d, S, near, far)
    assert torch.equal(ts.cpu(), t_ref) and torch.equal(pts.cpu(), p_ref)
    torch.manual_seed(5)
    rand = torch.rand(150, S)
    pts, ts = renderer_fp32.sample_uniform(o, d, S, near, far, rand=rand)
    p_ref, t_ref = O.sample_uniform(o, d, S, near, far, rand=rand)
    assert torch.equal(ts.cpu(), t_ref) and torch.equal(pts.cpu(), p_ref)


def test_sample_uniform_matches_golden_chain(renderer_fp32):
    g = golden("chain_w1")
    pts, ts = renderer_fp32.sample_uniform(g["rays_o"], g["rays_d"], 64, 0., 1.)
    assert np.array_equal(ts.cpu().numpy(), g["ts"]) and np.array_equal(pts.cpu().numpy(), g["pts_coarse"])


# ------------------------------------------------------------------ K5
def test_composite_vs_golden_and_oracle(renderer_fp32):
    g = golden("stages")
    rgb, depth, w, acc = renderer_fp32.composite(g["rgb_pts"], g["sigma"], g["ts_fine"])
    np.testing.assert_allclose(rgb.cpu().numpy(), g["rgb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(w.cpu().numpy(), g["weights"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(acc.cpu().numpy(), g["weights"].sum(-1), atol=2e-6, rtol=0)
    rgbw = renderer_fp32.composite(g["rgb_pts"], g["sigma"], g["ts_fine"], white_bkgd=True)[0]
    np.testing.assert_allclose(rgbw.cpu().numpy(), g["rgb_white"], atol=3e-6, rtol=0)
    # packed (r,g,b,sigma) input is the same function
    rs = torch.cat([torch.from_numpy(g["rgb_pts"]), torch.from_numpy(g["sigma"])[..., None]], -1)
    rgb2, depth2, w2, acc2 = renderer_fp32.composite(t_values=g["ts_fine"], rgbsigma=rs)
    assert torch.equal(rgb2, rgb) and torch.equal(w2, w) and torch.equal(depth2, depth)


@pytest.mark.parametrize("S", [64, 128, 50, 31])
def test_composite_shapes_noise_and_shared_ts(renderer_fp32, S):
    torch.manual_seed(S)
    n = 257
    sig = torch.randn(n, S) * 40.
    rgbp = torch.rand(n, S, 3)
    ts_row = torch.sort(torch.rand(S))[0]
    ts = ts_row.unsqueeze(0).expand(n, S)            # stride-0 view like utils.py:512
    noise = torch.randn(n, S)
    ref = O.alpha_composition(rgbp, sig, ts, noise=noise)
    got = renderer_fp32.composite(rgbp, sig, ts, noise=noise)
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), atol=3e-6, rtol=0)
    # weights are a distribution prefix: >= 0, sum <= 1 (+ rounding)
    assert (got[2] >= 0).all() and (got[3] <= 1 + 1e-5).all()


def test_composite_early_termination_is_exact(renderer_fp32):
    """opaque first samples drive the transmittance to exactly 0; the early-out must give the same bits."""
    n, S = 64, 128
    sig = torch.full((n, S), 1e6)
    sig[:, 100:] = -5.
    rgbp = torch.rand(n, S, 3)
    ts = torch.linspace(0, 1, S).unsqueeze(0).expand(n, S).contiguous()
    ref = O.alpha_composition(rgbp, sig, ts)
    got = renderer_fp32.composite(rgbp, sig, ts)
    np.testing.assert_allclose(got[2].cpu().numpy(), ref[2].numpy(), atol=1e-12, rtol=0)
    np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), atol=1e-6, rtol=0)


# ------------------------------------------------------------------ K6+K7
def test_sample_fine_bit_exact_vs_golden_stages(renderer_fp32):
    """adversarial weights (empty rays, one-hot, tiny, mass in the last bin): indices and ts_fine bit-exact
    against vectors from the reference's own sample_pdf/searchsorted/sort."""
    g = golden("stages")
    n = g["ts"].shape[0]
    pts, ts_f, smp, inds = renderer_fp32.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), g["ts"], g["weights_in"], 64, return_aux=True)
    assert np.array_equal(inds.cpu().numpy(), g["pdf_inds"])
    assert np.array_equal(ts_f.cpu().numpy(), g["ts_fine"])
    assert np.array_equal(pts.cpu().numpy()[..., 0], g["ts_fine"])      # o=0, d=1 -> pts == t


@pytest.mark.parametrize("kind", ["w1", "w0"])
def test_sample_fine_bit_exact_vs_golden_chain(renderer_fp32, kind):
    g = golden("chain_" + kind)
    pts, ts_f, smp, inds = renderer_fp32.sample_fine(g["rays_o"], g["rays_d"], g["ts"], g["weights_coarse"], 64, return_aux=True)
    assert np.array_equal(inds.cpu().numpy(), g["pdf_inds"])
    assert np.array_equal(ts_f.cpu().numpy(), g["ts_fine"])
    ref_pts = O.sample_fine(g["rays_o"], g["rays_d"], g["ts"], g["weights_coarse"], 64)[0]
    assert torch.equal(pts.cpu(), ref_pts)


@pytest.mark.parametrize("S,F", [(64, 64), (64, 128), (32, 32), (128, 128), (48, 17), (10, 5)])
def test_sample_fine_bit_exact_random(renderer_fp32, S, F):
    torch.manual_seed(S * 1000 + F)
    n = 3001
    ts = torch.sort(torch.rand(n, S), -1)[0]
    w = torch.rand(n, S) ** 6 * (torch.rand(n, 1) < 0.8)
    ref_pts, ref_ts, ref_smp, ref_inds, _ = O.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), ts, w, F, return_aux=True)
    pts, ts_f, smp, inds = renderer_fp32.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), ts, w, F, return_aux=True)
    assert torch.equal(inds.cpu(), ref_inds)
    assert torch.equal(smp.cpu(), ref_smp)
    assert torch.equal(ts_f.cpu(), ref_ts)
    # shared (stride-0) coarse ts row, as the render pipeline passes it
    row = torch.from_numpy(O.linspace_f32(0., 1., S))
    ref_ts2 = O.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), row.unsqueeze(0).expand(n, S), w, F)[1]
    ts2 = renderer_fp32.sample_fine(None, None, row, w, F, want_pts=False)[1]
    assert torch.equal(ts2.cpu(), ref_ts2)
    assert (ts2[:, 0] == 0).all() and (ts2[:, -1] == 1).all() and (ts2[:, 1:] >= ts2[:, :-1]).all()


@pytest.mark.gpu
def test_sample_fine_lean_64_kernel_equals_general_kernel_and_oracle(renderer_fp32):
    """The 64 + 64 ts_fine-only kernel (registers + shuffles, fixed-trip searches, occupancy-mask merge; sampling.cu) against the
    general kernel and the oracle, bit for bit: random and adversarial weights (all zero, one-hot, two spikes, tiny, saturated),
    per-ray sorted positions, the shared row, and position rows with duplicates (the merge's check-and-fall-back path)."""
    from tgtc_style_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(7)
    n, S, F = 4099, 64, 64
    w = torch.rand(n, S) ** 6 * (torch.rand(n, 1) < 0.8)
    w[0] = 0
    w[1] = 0; w[1, 17] = 1.0
    w[2] = 0; w[2, 1] = 0.5; w[2, 62] = 0.5
    w[3] = 1e-12
    w[4] = 1.0
    w[5] = 0; w[5, 31] = 1e-6; w[5, 32] = 1.0
    w[6] = torch.rand(S) * 1e-7
    ts_row = torch.from_numpy(O.linspace_f32(0., 1., S))
    ts_ray = torch.sort(torch.rand(n, S), -1)[0]
    ts_dup = ts_ray.clone()
    ts_dup[:, 10:14] = ts_dup[:, 10:11]            # runs of equal positions
    ts_dup[::3, 40:64] = ts_dup[::3, 40:41]
    for name, ts in (("row", ts_row), ("ray", ts_ray), ("dup", ts_dup)):
        ts_full = ts.unsqueeze(0).expand(n, S) if ts.dim() == 1 else ts
        ref = O.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), ts_full, w, F)[1]
        lean = renderer_fp32.sample_fine(None, None, ts, w, F, want_pts=False)[1].cpu()
        lib.tgtc_debug_sample_fine_general(1)
        try:
            general = renderer_fp32.sample_fine(None, None, ts, w, F, want_pts=False)[1].cpu()
        finally:
            lib.tgtc_debug_sample_fine_general(0)
        assert torch.equal(general, ref), name
        assert torch.equal(lean, ref), name
