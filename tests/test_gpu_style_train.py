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
    # (g + g_coarse) + g_fine against 2 * (g_coarse + g_fine): equal up to one rounding of the largest term
    assert (bw2["grads"] - 2 * bw["grads"]).abs().max().item() <= 4e-7 * bw["grads"].abs().max().item()
    assert torch.equal(bw2["d_latents"], bw["d_latents"])
    # a fresh forward + backward of the same batch reproduces everything bit for bit (ordered reductions, no atomics)
    fw3 = r.style_train_forward(ro, rd, lat, rand=rand)
    bw3 = r.style_train_backward(fw3["state"], g_c, g_f)
    torch.cuda.synchronize()
    assert torch.equal(fw3["rgb_fine"], fw["rgb_fine"]) and torch.equal(bw3["grads"], bw["grads"]) and torch.equal(bw3["d_latents"], bw["d_latents"])


def test_style_train_rejects_bad_arguments(renderer_bf16):
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, _, _, _ = _inputs(64)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    with pytest.raises(ValueError):
        r.style_train_forward(ro, rd, lat[:10])
    with pytest.raises(Exception):
        r.style_train_forward(ro, rd, lat, n_samples=32)
    # a backward needs the workspace of a matching forward
    import tgtc_style_b200 as T
    fw = r.style_train_forward(ro, rd, lat)
    g = torch.zeros(64, 3)
    r2 = T.NerfRenderer(device=r.device, mode="bf16")       # a context that never ran a forward
    r2.set_weights(wc, wf)
    r2.set_style_weights(cs, ws)
    with pytest.raises(Exception, match="no tgtc_style_train_forward"):
        r2.style_train_backward(fw["state"], g, g)
    r2.close()
    bad2 = dict(fw["state"])
    bad2["rand"] = torch.zeros(64, 64)
    with pytest.raises(Exception, match="different batch"):
        r.style_train_backward(bad2, g, g)
    r.style_train_backward(fw["state"], g, g)


def _batch(ro, rd, idx, style_num, frame_num, g, origin=False):
    n = len(idx)
    idx = idx.numpy()
    b = {"rays_o": torch.from_numpy(ro[idx]), "rays_d": torch.from_numpy(rd[idx]), "rgb_gt": torch.rand(n, 3, generator=g),
         "style_id": torch.randint(0, style_num, (n,), generator=g), "frame_id": torch.randint(0, frame_num, (n,), generator=g),
         "rand": torch.rand(n, 64, generator=g)}
    if origin:
        b["rgb_origin"] = torch.rand(n, 3, generator=g)
    return b


def test_style_trainer_step_vs_oracle(renderer_bf16):
    """Two StyleTrainer iterations (train_tgtcs.py:354-495): the second one carries the coherence term against the first one's
    maps.  Losses, style-module gradients and the gradient of the latent table against torch.autograd through the oracle."""
    import tgtc_style_b200 as T
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, _, _, _, _ = _inputs(8)
    ro_all, rd_all = small_rays()
    r.set_weights(wc, wf)
    g = torch.Generator().manual_seed(21)
    style_num, frame_num, n = 2, 5, 128
    table = torch.randn(style_num, frame_num, 32, generator=g) * 0.5
    mu, logvar = torch.randn(style_num, 32, generator=g) * 0.3, torch.randn(style_num, 32, generator=g) * 0.2
    dev = r.device
    lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev), dataset_type="llff")
    tr = T.StyleTrainer(r, cs, ws, lat, lr=5e-4, loss_coh_lambda=1e2, frame_num=frame_num)
    to_dev = lambda b: {k: v.to(dev) for k, v in b.items()}
    perm = torch.randperm(ro_all.shape[0], generator=g)
    b1, c1 = _batch(ro_all, rd_all, perm[:n], style_num, frame_num, g), _batch(ro_all, rd_all, perm[n:2 * n], style_num, frame_num, g, True)
    out1 = tr.step(to_dev(b1), to_dev(c1))
    assert out1["loss_coh"].item() == 0.0 and tr.prev is not None          # cnt == 0: no coherence term yet (train_tgtcs.py:400)
    # second iteration: oracle on the trainer's current parameters / latent table / previous maps
    sdc, sdw = (dict((k, v.cpu().clone()) for k, v in d.items()) for d in tr.state_dicts())
    tab = lat.latents.detach().cpu().clone()
    prev = tuple(t.cpu() for t in tr.prev)
    b2, c2 = _batch(ro_all, rd_all, perm[2 * n:3 * n], style_num, frame_num, g), _batch(ro_all, rd_all, perm[3 * n:4 * n], style_num, frame_num, g, True)
    losses, gcs, gws, gtab = O.style_train_step_reference(wc, wf, sdc, sdw, tab, mu, logvar, b2, c2, prev, frame_num, loss_coh_lambda=1e2)
    out2 = tr.step(to_dev(b2), to_dev(c2))
    torch.cuda.synchronize()
    for k in ("loss_rgb", "loss_logp", "loss_coh", "loss"):
        a, b = out2[k].item(), losses[k].item()
        print("%-9s ours %.6f oracle %.6f" % (k, a, b))
        assert abs(a - b) <= 2e-2 * max(abs(b), 1e-3), (k, a, b)
    assert out2["loss_coh"].item() > 0.0
    ours_c, ours_w = r.style_grad_views(tr.grads)
    for ref, ours, tag in ((gcs, ours_c, "concat"), (gws, ours_w, "wild")):
        for k, gr in ref.items():
            a, b = ours[k].cpu().double().flatten(), gr.double().flatten()
            rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
            cos = (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
            assert rel <= 0.15 and cos >= 0.99, (tag, k, rel, cos)
    a, b = lat.latents.grad.cpu().double().flatten(), gtab.double().flatten()
    rel, cos = ((a - b).norm() / b.norm()).item(), (torch.dot(a, b) / (a.norm() * b.norm())).item()
    print("latent table grad rel %.3e cos %.5f" % (rel, cos))
    assert rel <= 0.1 and cos >= 0.995
    # the step moved the parameters and the table
    assert not torch.equal(lat.latents.detach().cpu(), tab)
    assert not torch.equal(tr.state_dicts()[1]["layers.7.weight"].cpu(), sdw["layers.7.weight"])


def test_style_train_forward_vs_reference_golden(renderer_bf16):
    """The training forward on the golden batch of oracle/make_golden_style_train.py (the imported reference's own modules,
    perturb=True replayed through the stored uniforms, per-ray latents from StyleLatents_variational)."""
    import tgtc_style_b200 as T
    from helpers import golden
    g = golden("style_train")
    r = renderer_bf16
    (wc, wf, cs, ws), _, _, _, _, _, _ = _inputs(8)
    ro, rd = small_rays()
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    dev = r.device
    lat = T.StyleLatents(torch.from_numpy(g["table"]).to(dev), torch.from_numpy(g["mu"]).to(dev), torch.from_numpy(g["logvar"]).to(dev))
    for tag, kc, kf in (("b1", "rgb_coarse", "rgb_fine"), ("b2", "coh_rgb_coarse", "coh_rgb_fine")):
        idx = g[tag + "/ray_index"]
        l1 = lat(torch.from_numpy(g[tag + "/style_id"]).to(dev), torch.from_numpy(g[tag + "/frame_id"]).to(dev)).detach()
        out = r.style_train_forward(ro[idx], rd[idx], l1, rand=torch.from_numpy(g[tag + "/rand"]))
        torch.cuda.synchronize()
        ec = np.abs(out["rgb_coarse"].cpu().numpy() - g[kc])
        ef = np.abs(out["rgb_fine"].cpu().numpy() - g[kf])
        print("%s: coarse max %.3e mean %.3e, fine max %.3e mean %.3e" % (tag, ec.max(), ec.mean(), ef.max(), ef.mean()))
        assert ec.mean() <= 3e-3 and ec.max() <= 3e-2
        assert ef.mean() <= 5e-3
    # minus_logp of the host latent model against the reference's (models.py:531-537)
    lp = 0.1 * lat.minus_logp(torch.from_numpy(g["b1/style_id"]).to(dev), torch.from_numpy(g["b1/frame_id"]).to(dev))
    assert abs(lp.item() - float(g["loss_logp"])) <= 1e-5 * max(1.0, float(g["loss_logp"]))


@pytest.mark.parametrize("n", [1, 3, 130])
def test_style_train_ragged_batches(renderer_bf16, n):
    """ray counts that leave half-filled / padding tiles in both passes (coarse tile = 2 rays, fine tile = 1 ray, tiles are
    handed out in quads): forward, parameter gradients and latent gradients still match autograd"""
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, rand, g_c, g_f = _inputs(n, seed=17 + n)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    ref_c, ref_f, gcs, gws, dlat_ref = O.style_train_forward_backward(wc, wf, cs, ws, ro, rd, lat, g_c, g_f, rand=rand)
    fw = r.style_train_forward(ro, rd, lat, rand=rand)
    bw = r.style_train_backward(fw["state"], g_c, g_f)
    torch.cuda.synchronize()
    assert (fw["rgb_coarse"].cpu() - ref_c).abs().max().item() <= 3e-2
    assert (fw["rgb_fine"].cpu() - ref_f).abs().max().item() <= 5e-2
    ours_c, ours_w = r.style_grad_views(bw["grads"])
    for ref, ours, tag in ((gcs, ours_c, "concat"), (gws, ours_w, "wild")):
        for k, gr in ref.items():
            a, b = ours[k].cpu().double().flatten(), gr.double().flatten()
            assert torch.isfinite(a).all(), (tag, k)
            rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
            assert rel <= 0.2, (n, tag, k, rel)
    a, b = bw["d_latents"].cpu().double().flatten(), dlat_ref.double().flatten()
    assert ((a - b).norm() / b.norm()).item() <= 0.1


def test_style_wgrad_work_splits_agree(renderer_bf16):
    """the two work splits of style_wgrad_kernel (all jobs per CTA for large batches / one job per CTA for small ones) give the
    same gradients up to summation order"""
    import ctypes
    from tgtc_style_b200 import _lib
    r = renderer_bf16
    (wc, wf, cs, ws), ro, rd, lat, rand, g_c, g_f = _inputs(200, seed=31)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    lib = _lib.load()
    lib.tgtc_debug_style_wgrad_mode.argtypes = [ctypes.c_int]
    fw = r.style_train_forward(ro, rd, lat, rand=rand)
    res = []
    try:
        for mode in (0, 1):
            lib.tgtc_debug_style_wgrad_mode(mode)
            bw = r.style_train_backward(fw["state"], g_c, g_f)
            torch.cuda.synchronize()
            res.append((bw["grads"].clone(), bw["d_latents"].clone()))
    finally:
        lib.tgtc_debug_style_wgrad_mode(-1)
    (ga, la), (gb, lb) = res
    assert ga.abs().max().item() > 0
    assert ((ga - gb).norm() / ga.norm()).item() <= 1e-5
    assert torch.equal(la, lb)          # the R rows do not depend on the split


def test_style_train_seeded_equals_replay(renderer_bf16):
    """tgtc_style_train_forward/backward_seeded (jitter + sigma noise drawn in-kernel) == the replay entries fed the same streams"""
    r = renderer_bf16
    n = 96
    (wc, wf, cs, ws), ro, rd, lat, _, g_c, g_f = _inputs(n, seed=41)
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    seed, std = 987654321, 0.5
    fa = r.style_train_forward(ro, rd, lat, seed=seed, perturb=True, sigma_noise_std=std)
    ba = r.style_train_backward(fa["state"], g_c, g_f)
    rand = r.philox_fill(seed, 0, n * 64).view(n, 64)
    nzc = r.philox_fill(seed, 1, n * 64, normal=True, std=std).view(n, 64)
    nzf = r.philox_fill(seed, 2, n * 128, normal=True, std=std).view(n, 128)
    fb = r.style_train_forward(ro, rd, lat, rand=rand, noise_coarse=nzc, noise_fine=nzf)
    bb = r.style_train_backward(fb["state"], g_c, g_f)
    torch.cuda.synchronize()
    assert torch.equal(fa["rgb_coarse"], fb["rgb_coarse"]) and torch.equal(fa["rgb_fine"], fb["rgb_fine"])
    assert torch.equal(ba["grads"], bb["grads"]) and torch.equal(ba["d_latents"], bb["d_latents"])
    fc = r.style_train_forward(ro, rd, lat, seed=seed + 1, perturb=True, sigma_noise_std=std)
    assert not torch.equal(fc["rgb_fine"], fa["rgb_fine"])


def test_style_trainer_fused_losses_equal_torch_losses(renderer_bf16):
    """tgtc_style_loss_sums / tgtc_style_loss_grads against the same losses written with torch ops + autograd: two iterations
    (the second with the coherence term), same batches, same jitter"""
    import tgtc_style_b200 as T
    r = renderer_bf16
    (wc, wf, cs, ws), _, _, _, _, _, _ = _inputs(8)
    ro_all, rd_all = small_rays()
    r.set_weights(wc, wf)
    dev = r.device
    g = torch.Generator().manual_seed(77)
    style_num, frame_num, n = 2, 5, 96
    table = torch.randn(style_num, frame_num, 32, generator=g) * 0.5
    mu, logvar = torch.randn(style_num, 32, generator=g) * 0.3, torch.randn(style_num, 32, generator=g) * 0.2
    perm = torch.randperm(ro_all.shape[0], generator=g)
    its = [(_batch(ro_all, rd_all, perm[(2 * i) * n:(2 * i + 1) * n], style_num, frame_num, g),
            _batch(ro_all, rd_all, perm[(2 * i + 1) * n:(2 * i + 2) * n], style_num, frame_num, g, True)) for i in range(2)]
    # a zero row in the originals: cos(0, .) = 0 and its gradient is the zero subgradient, as in torch
    its[1][1]["rgb_origin"][3] = 0.0
    res = []
    for fused in (True, False):
        lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev))
        tr = T.StyleTrainer(r, cs, ws, lat, frame_num=frame_num, fused_losses=fused)
        assert tr.fused_losses == fused
        out = None
        for b, c in its:
            out = tr.step({k: v.to(dev) for k, v in b.items()}, {k: v.to(dev) for k, v in c.items()})
        torch.cuda.synchronize()
        res.append((tr.grads.clone(), lat.latents.grad.clone(), {k: v.item() for k, v in out.items()}))
    (ga, ta, la), (gb, tb, lb) = res
    assert la["loss_coh"] > 0
    for k in la:
        assert abs(la[k] - lb[k]) <= 1e-5 * max(1.0, abs(lb[k])), (k, la[k], lb[k])
    assert ((ga - gb).norm() / gb.norm()).item() <= 2e-4
    assert ((ta - tb).norm() / tb.norm()).item() <= 2e-4


@pytest.mark.parametrize("fused", [True, False])
def test_latent_table_gradient_ignores_the_coherence_term(renderer_bf16, fused):
    """latents_model_1.optimize(loss) backpropagates loss = loss_rgb + loss_logp (train_tgtcs.py:481, :495; models.py:544-549):
    the coherence term reaches the two style modules (loss_for_style, :482-493) but never the latent table -- its gradient must
    not depend on loss_coh_lambda, while the style gradient does."""
    import tgtc_style_b200 as T
    r = renderer_bf16
    (wc, wf, cs, ws), _, _, _, _, _, _ = _inputs(8)
    ro_all, rd_all = small_rays()
    r.set_weights(wc, wf)
    dev = r.device
    g = torch.Generator().manual_seed(78)
    style_num, frame_num, n = 2, 5, 64
    table = torch.randn(style_num, frame_num, 32, generator=g) * 0.5
    mu, logvar = torch.randn(style_num, 32, generator=g) * 0.3, torch.randn(style_num, 32, generator=g) * 0.2
    perm = torch.randperm(ro_all.shape[0], generator=g)
    its = [(_batch(ro_all, rd_all, perm[(2 * i) * n:(2 * i + 1) * n], style_num, frame_num, g),
            _batch(ro_all, rd_all, perm[(2 * i + 1) * n:(2 * i + 2) * n], style_num, frame_num, g, True)) for i in range(2)]
    res = []
    for lam in (1e2, 0.0):
        lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev))
        # lr = 0: the first iteration leaves every parameter where it was, so both runs reach iteration 2 in the same state
        tr = T.StyleTrainer(r, cs, ws, lat, lr=0.0, frame_num=frame_num, loss_coh_lambda=lam, fused_losses=fused)
        lat.lr = 0.0
        if lat.opt is not None:
            for gp in lat.opt.param_groups:
                gp["lr"] = 0.0
        for b, c in its:
            out = tr.step({k: v.to(dev) for k, v in b.items()}, {k: v.to(dev) for k, v in c.items()})
        torch.cuda.synchronize()
        assert out["loss_coh"].item() > 0
        res.append((tr.grads.clone(), lat.latents.grad.clone()))
    (g_with, t_with), (g_without, t_without) = res
    assert ((t_with - t_without).norm() / t_without.norm()).item() <= 1e-5
    assert ((g_with - g_without).norm() / g_without.norm()).item() >= 1e-3
