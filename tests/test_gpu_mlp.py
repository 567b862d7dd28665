"""GPU parity of K3+K4 (positional encoding + MLP) through the C ABI: the fp32 CUDA-core kernel against the
oracle, and the bf16 tcgen05 kernel layer by layer against a bf16-operand emulation and the oracle."""
import numpy as np
import pytest
import torch

import render_oracle as O
import tgtc_style_b200 as T
from helpers import emulate_bf16_forward, golden, small_rays, weights

pytestmark = pytest.mark.gpu


def _inputs(n, S, seed=0):
    ro, rd = small_rays()
    rng = np.random.RandomState(seed)
    sel = rng.choice(ro.shape[0], n, replace=False)
    o, d = torch.from_numpy(ro[sel]), torch.from_numpy(rd[sel])
    if S == 64:
        pts, ts = O.sample_uniform(o, d, 64, 0., 1.)
        ts = ts.contiguous()
    else:
        ts = torch.sort(torch.rand(n, S, generator=torch.Generator().manual_seed(seed)), -1)[0]
        pts = o.unsqueeze(1) + ts.unsqueeze(-1) * d.unsqueeze(1)
    return o, d, pts, ts


# ------------------------------------------------------------------ fp32 kernel
@pytest.mark.parametrize("kind", ["w1", "w0"])
@pytest.mark.parametrize("n,S", [(96, 64), (37, 128), (5, 50)])
def test_fp32_forward_matches_oracle(renderer_fp32, kind, n, S):
    wc, wf = weights(kind)
    renderer_fp32.set_weights(wc, wf)
    o, d, pts, ts = _inputs(n, S)
    dirs = d.unsqueeze(1).expand(n, S, 3)                 # stride-0 view (rendering.py:30)
    for which, sd in ((T.NET_COARSE, wc), (T.NET_FINE, wf)):
        ref = O.nerf_forward(sd, pts, dirs)
        got = renderer_fp32.nerf_forward(which, pts, dirs, want_features=True)
        np.testing.assert_allclose(got["rgb"].cpu().numpy(), ref["rgb"].numpy(), atol=2e-5, rtol=0)
        s_ref = ref["sigma"].numpy()
        np.testing.assert_allclose(got["sigma"].cpu().numpy(), s_ref, atol=1e-4 * max(1.0, np.abs(s_ref).max()), rtol=0)
        np.testing.assert_allclose(got["base_remap"].cpu().numpy(), ref["base_remap"].numpy(), atol=2e-5, rtol=1e-4)
        np.testing.assert_allclose(got["pts"].cpu().numpy(), ref["pts"].numpy(), atol=1e-6, rtol=0)
        np.testing.assert_allclose(got["dirs"].cpu().numpy(), ref["dirs"].numpy(), atol=1e-6, rtol=0)
    # fused ray-mode entry (pts formed in-kernel) equals the explicit-points entry bit for bit
    rs = renderer_fp32.nerf_forward_rays(T.NET_FINE, o, d, ts, S, 0., 1., mode="fp32")
    got = renderer_fp32.nerf_forward(T.NET_FINE, pts, dirs, want_features=False)
    assert torch.equal(rs[..., :3], got["rgb"]) and torch.equal(rs[..., 3], got["sigma"])


def test_fp32_forward_per_sample_dirs(renderer_fp32):
    """general path: view directions that differ per sample (not the stride-0 expand)."""
    wc, wf = weights("w1")
    renderer_fp32.set_weights(wc, wf)
    torch.manual_seed(1)
    pts = torch.rand(11, 64, 3) * 2 - 1
    dirs = torch.randn(11, 64, 3)
    ref = O.nerf_forward(wc, pts, dirs)
    got = renderer_fp32.nerf_forward(T.NET_COARSE, pts, dirs, want_features=False)
    np.testing.assert_allclose(got["rgb"].cpu().numpy(), ref["rgb"].numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(got["sigma"].cpu().numpy(), ref["sigma"].numpy(), atol=5e-3, rtol=1e-4)


# ------------------------------------------------------------------ bf16 tcgen05 kernel
@pytest.mark.parametrize("layers", [1, 2, 5, 6, 8, 9, 10])
def test_tc_accumulators_layer_by_layer(renderer_bf16, layers):
    """debug hook: raw fp32 TMEM accumulator of GEMM layer `layers` (after running all earlier layers on
    chip) against the bf16-operand emulation.  Catches descriptor / swizzle / pipeline bugs layer by layer."""
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    n, S = 150, 64          # 75 tiles: most CTAs own one tile, exercising single-slot iterations too
    o, d, pts, ts = _inputs(n, S, seed=3)
    emu = emulate_bf16_forward(wc, pts.reshape(-1, 3), d, S)
    acc = renderer_bf16.debug_tc_layers(T.NET_COARSE, o, d, None, S, 0., 1., layers).cpu()
    ref = emu["accs"][layers - 1]
    N = ref.shape[1]
    err = (acc[:, :N] - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp32 accumulation-order noise + rare bf16 rounding flips of upstream activations
    assert err <= 3e-2 * scale, (layers, err, scale)
    assert (acc[:, :N] - ref).abs().mean().item() <= 2e-3 * scale


@pytest.mark.parametrize("n,S", [(300, 64), (301, 64), (77, 128), (1, 64), (3, 256)])
def test_tc_forward_vs_emulation_and_oracle(renderer_bf16, n, S):
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    o, d, pts, ts = _inputs(n, S, seed=4)
    rs = renderer_bf16.nerf_forward_rays(T.NET_FINE, o, d, None if S == 64 else ts, S, 0., 1.).cpu()
    emu = emulate_bf16_forward(wf, pts.reshape(-1, 3), d, S)
    rgb_e, sig_e = emu["rgb"].reshape(n, S, 3), emu["sigma"].reshape(n, S)
    assert (rs[..., :3] - rgb_e).abs().max().item() <= 5e-3
    assert (rs[..., 3] - sig_e).abs().max().item() <= 2e-2 * sig_e.abs().max().item()
    ref = O.nerf_forward(wf, pts, d.unsqueeze(1).expand(n, S, 3))
    # per-sample accuracy of the bf16 path against the fp32 oracle (|drgb| ~1.5e-4, |dsigma|/|sigma| ~1%)
    assert (rs[..., :3] - ref["rgb"]).abs().max().item() <= 1e-2
    # sigma: bf16 operands cost ~1% of the sigma scale; the kernel must be as close to the oracle as the
    # bf16-operand emulation is (same arithmetic, different summation order)
    scale = ref["sigma"].abs().max().item()
    err_k = (rs[..., 3] - ref["sigma"]).abs()
    err_e = (sig_e - ref["sigma"]).abs()
    assert err_k.median().item() <= 1e-2 * scale and err_k.max().item() <= 5e-2 * scale
    assert err_k.max().item() <= 1.25 * err_e.max().item() + 1e-3 * scale
    # explicit-points entry (the model_forward shim) gives the same numbers as the fused ray entry
    got = renderer_bf16.nerf_forward(T.NET_FINE, pts, d.unsqueeze(1).expand(n, S, 3), want_features=False, mode="bf16")
    assert torch.allclose(got["rgb"].cpu(), rs[..., :3], atol=1e-6) and torch.allclose(got["sigma"].cpu(), rs[..., 3], atol=1e-4)


def test_tc_unsupported_shapes_fail_loudly(renderer_bf16):
    wc, wf = weights("w1")
    renderer_bf16.set_weights(wc, wf)
    o, d, pts, ts = _inputs(4, 64)
    with pytest.raises(T.TgtcError):
        renderer_bf16.nerf_forward_rays(T.NET_COARSE, o, d, None, 50, 0., 1.)     # S=50 not tileable
    with pytest.raises(T.TgtcError):
        renderer_bf16.nerf_forward(T.NET_COARSE, torch.rand(4, 64, 3), torch.rand(4, 64, 3), want_features=False, mode="bf16")
