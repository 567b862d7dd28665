"""Style_train on the GPU (train_tgtcs.py:311-495): forward with per-ray latents + stash, backward into the two
style modules and the latents, against torch.autograd through the CPU oracle chain."""
import numpy as np
import pytest
import torch

import render_oracle as O
from helpers import small_rays, weights

pytestmark = pytest.mark.gpu


def _inputs(n, seed=5, n_latents=3):
    ro, rd = small_rays()
    rng = np.random.RandomState(seed)
    sel = rng.permutation(ro.shape[0])[:n]
    w0c, w0f = weights("w0")
    probe = np.arange(0, ro.shape[0], 743)
    # smooth, non-degenerate density (see test_gpu_backward._train_inputs for why sigma is kept away from 0)
    wc = O.recalibrate_sigma(w0c, ro[probe], rd[probe], gain=4.0, shift=10.0)
    wf = O.recalibrate_sigma(w0f, ro[probe], rd[probe], gain=4.0, shift=10.0)
    cs, ws = O.init_style_like_reference(1)
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(n_latents, 32, generator=g) * 0.7          # a shuffled batch: every ray has its own (style, frame)
    lat = table[torch.randint(0, n_latents, (n,), generator=g)]
    rand = torch.rand(n, 64, generator=g)
    g_c = torch.randn(n, 3, generator=g) / n
    g_f = torch.randn(n, 3, generator=g) / n
    return (wc, wf, cs, ws), ro[sel], rd[sel], lat, rand, g_c, g_f


def test_style_train_forward_matches_render_style(renderer_bf16):
    """all rays share one latent, perturb off: the training forward must reproduce the inference path"""
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, _, _, _ = _inputs(300)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    one = lat[:1].expand(300, 32).contiguous()
    a = r.render_style(ro, rd, one[0], extras=True)
    b = r.style_train_forward(ro, rd, one)
    torch.cuda.synchronize()
    assert (a["rgb_coarse"] - b["rgb_coarse"]).abs().max().item() <= 2e-3
    assert (a["rgb"] - b["rgb_fine"]).abs().mean().item() <= 2e-3


def test_style_train_forward_per_ray_latents_vs_oracle(renderer_bf16):
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, rand, g_c, g_f = _inputs(256)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    ref_c, ref_f = O.style_train_forward_backward(wc, wf, cs, ws, ro, rd, lat, g_c, g_f, rand=rand)[:2]
    out = r.style_train_forward(ro, rd, lat, rand=rand)
    torch.cuda.synchronize()
    ec = (out["rgb_coarse"].cpu() - ref_c).abs()
    ef = (out["rgb_fine"].cpu() - ref_f).abs()
    print("style-train fwd: coarse max %.3e mean %.3e, fine max %.3e mean %.3e" % (ec.max(), ec.mean(), ef.max(), ef.mean()))
    assert ec.mean().item() <= 3e-3 and ec.max().item() <= 3e-2
    assert ef.mean().item() <= 5e-3


def test_style_train_gradients_vs_autograd(renderer_bf16):
    """bf16 tcgen05 forward+backward against fp32 torch.autograd (the reference's graph, train_tgtcs.py:404-495)."""
    n = 256
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, rand, g_c, g_f = _inputs(n)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    _, _, gcs, gws, dlat_ref = O.style_train_forward_backward(wc, wf, cs, ws, ro, rd, lat, g_c, g_f, rand=rand)
    fw = r.style_train_forward(ro, rd, lat, rand=rand)
    bw = r.style_train_backward(fw["state"], g_c, g_f)
    torch.cuda.synchronize()
    ours_c, ours_w = r.style_grad_views(bw["grads"])
    worst = 0.0
    for ref, ours, tag in ((gcs, ours_c, "concat"), (gws, ours_w, "wild")):
        for k, gr in ref.items():
            a, b = ours[k].cpu().double().flatten(), gr.double().flatten()
            assert torch.isfinite(a).all(), (tag, k)
            rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
            cos = (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
            print("%s %-18s rel %.3e cos %.5f |g| %.3e" % (tag, k, rel, cos, b.norm().item()))
            worst = max(worst, rel)
            assert rel <= 0.12 and cos >= 0.995, (tag, k, rel, cos)
    a, b = bw["d_latents"].cpu().double().flatten(), dlat_ref.double().flatten()
    rel = ((a - b).norm() / b.norm()).item()
    cos = (torch.dot(a, b) / (a.norm() * b.norm())).item()
    print("d_latents rel %.3e cos %.5f" % (rel, cos))
    assert rel <= 0.12 and cos >= 0.995
    # accumulate=True adds to an existing buffer; a second backward of the same state is reproducible bit for bit
    g2 = bw["grads"].clone()
    bw2 = r.style_train_backward(fw["state"], g_c, g_f, grads=g2, accumulate=True)
    torch.cuda.synchronize()
    assert torch.allclose(bw2["grads"], 2 * bw["grads"], rtol=1e-6, atol=1e-12)
    assert torch.equal(bw2["d_latents"], bw["d_latents"])


def test_style_train_rejects_bad_arguments(renderer_bf16):
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, _, _, _ = _inputs(64)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    with pytest.raises(ValueError):
        r.style_train_forward(ro, rd, lat[:10])
    with pytest.raises(Exception):
        r.style_train_forward(ro, rd, lat, n_samples=32)
