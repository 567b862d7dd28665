"""Backward / training-step parity: our CUDA gradients against torch.autograd through the CPU
oracle chain (the reference's own autograd graph, train_tgtcs.py:236-255), perturb=0, noise=0."""
import numpy as np
import pytest
import torch

import render_oracle as O
from helpers import small_rays, weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,white", [(64, False), (128, False), (128, True), (50, False)])
def test_composite_backward_vs_autograd(renderer_fp32, S, white):
    torch.manual_seed(S)
    n = 193
    sig = (torch.randn(n, S) * 20.).requires_grad_(True)
    rgbp = torch.rand(n, S, 3).requires_grad_(True)
    ts = torch.sort(torch.rand(n, S), -1)[0]
    rgb, depth, w = O.alpha_composition(rgbp, sig, ts, white_bkgd=white)[:3]
    gt = torch.rand(n, 3)
    g_depth = torch.randn(n) * 0.1
    loss = ((rgb - gt) ** 2).mean() + (depth * g_depth).sum()
    loss.backward()
    g_rgb = (2.0 * (rgb.detach() - gt) / (3 * n))
    rs = torch.cat([rgbp.detach(), sig.detach()[..., None]], -1)
    d = renderer_fp32.composite_backward(rs, ts, g_rgb, g_depth=g_depth, white_bkgd=white).cpu()
    scale = rgbp.grad.abs().max().item()
    np.testing.assert_allclose(d[..., :3].numpy(), rgbp.grad.numpy(), atol=2e-6 * max(scale, 1.0) + 1e-9, rtol=1e-4)
    gs = sig.grad
    err = (d[..., 3] - gs).abs()
    tol = 1e-4 * gs.abs().max().item() + 1e-9
    assert err.max().item() <= tol, (err.max().item(), tol)


def test_composite_backward_saturated_rays(renderer_fp32):
    """alpha == 1 inside the ray (factor 1e-10) and the 1e10 last interval: same gradients as autograd."""
    n, S = 32, 64
    sig = torch.full((n, S), -3.0)
    sig[:, 20] = 1e6
    sig[:, 40:] = 7.0
    sig = sig.requires_grad_(True)
    rgbp = torch.rand(n, S, 3).requires_grad_(True)
    ts = torch.linspace(0, 1, S).unsqueeze(0).expand(n, S).contiguous()
    rgb = O.alpha_composition(rgbp, sig, ts)[0]
    gt = torch.rand(n, 3)
    ((rgb - gt) ** 2).mean().backward()
    g_rgb = 2.0 * (rgb.detach() - gt) / (3 * n)
    d = renderer_fp32.composite_backward(torch.cat([rgbp.detach(), sig.detach()[..., None]], -1), ts, g_rgb).cpu()
    assert torch.isfinite(d).all()
    np.testing.assert_allclose(d[..., :3].numpy(), rgbp.grad.numpy(), atol=1e-8, rtol=1e-4)
    np.testing.assert_allclose(d[..., 3].numpy(), sig.grad.numpy(), atol=1e-6 * sig.grad.abs().max().item() + 1e-12, rtol=1e-3)
