"""CPU: the oracle (oracle/render_oracle.py) against the golden vectors produced by the REAL reference
(oracle/make_golden.py) and, when /root/reference exists, against the reference itself run live."""
import hashlib

import numpy as np
import pytest
import torch

import ref_import
import render_oracle as O
from helpers import FERN_FULL, FERN_SMALL, golden, small_rays, weights


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_weights_w0_identical_to_reference():
    g = golden("weights_w0_sha")
    wc, wf = O.init_linear_like_reference(0)
    for tag, sd in (("coarse", wc), ("fine", wf)):
        for k, v in sd.items():
            assert sha(v.numpy()) == str(g[tag + "/" + k]), k


def test_weights_w1_recalibration_identical():
    g = golden("chain_w1")
    wc, wf = weights("w1")
    assert sha(wc["net.sigma_layer.weight"].numpy()) == str(g["sha_sigma_w_coarse"])
    assert sha(wf["net.sigma_layer.weight"].numpy()) == str(g["sha_sigma_w_fine"])


@pytest.mark.parametrize("tag,hwf", [("small_identity", FERN_SMALL), ("full_identity", FERN_FULL), ("full_spiral17", FERN_FULL)])
def test_rays_bit_exact(tag, hwf):
    g = golden("rays")
    H, W, f = hwf
    ro, rd = O.make_rays(H, W, f, g[tag + "_c2w"])
    assert sha(ro) == str(g[tag + "_sha_o"]) and sha(rd) == str(g[tag + "_sha_d"])
    idx = g[tag + "_idx"]
    assert np.array_equal(ro[idx], g[tag + "_o"]) and np.array_equal(rd[idx], g[tag + "_d"])
    # NDC invariants the survey probes: o_z == -1, d_z == 2 exactly
    assert np.all(ro[:, 2] == -1.0) and np.all(rd[:, 2] == 2.0)


def test_linspace_matches_torch():
    for steps in (2, 3, 16, 63, 64, 65, 128, 192):
        assert np.array_equal(O.linspace_f32(0., 1., steps), torch.linspace(0., 1., steps).numpy())


def test_row_sum_and_cumsum_match_torch_here():
    """the explicit ATen-order restatements equal torch on this host (8-lane vector order)."""
    torch.manual_seed(3)
    for n in (62, 30, 126, 8, 40, 33):
        w = torch.rand(4096, n) * torch.rand(4096, 1) + 1e-5
        assert np.array_equal(O.row_sum_f32(w.numpy()), torch.sum(w, -1).numpy()), n
        p = w / torch.sum(w, -1, keepdim=True)
        assert np.array_equal(O.cumsum_f32(p.numpy()), torch.cumsum(p, -1).numpy()), n


def test_stage_sampling_bit_exact_vs_golden():
    """sample_pdf / sort on adversarial weights: searchsorted indices and ts_fine are bit-exact."""
    g = golden("stages")
    ts, w = torch.from_numpy(g["ts"]), torch.from_numpy(g["weights_in"])
    n = ts.shape[0]
    _, ts_f, _, inds, _ = O.sample_fine(torch.zeros(n, 3), torch.ones(n, 3), ts, w, 64, return_aux=True)
    assert np.array_equal(inds.numpy(), g["pdf_inds"])
    assert np.array_equal(ts_f.numpy(), g["ts_fine"])
    assert np.all(np.diff(ts_f.numpy(), axis=-1) >= 0)


def test_stage_compositing_vs_golden():
    g = golden("stages")
    rgb, depth, w, acc = O.alpha_composition(torch.from_numpy(g["rgb_pts"]), torch.from_numpy(g["sigma"]), torch.from_numpy(g["ts_fine"]))
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(depth.numpy(), g["depth"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(w.numpy(), g["weights"], atol=1e-6, rtol=0)
    rgbw = O.alpha_composition(torch.from_numpy(g["rgb_pts"]), torch.from_numpy(g["sigma"]), torch.from_numpy(g["ts_fine"]), white_bkgd=True)[0]
    np.testing.assert_allclose(rgbw.numpy(), g["rgb_white"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("kind", ["w1", "w0"])
def test_chain_vs_golden(kind):
    """whole chain (rendering.py:27-51).  Integer/bit-deciding stages are compared teacher-forced and bit-exact;
    MLP outputs within 2e-5 (CPU GEMM kernels differ between hosts)."""
    g = golden("chain_" + kind)
    wc, wf = weights(kind)
    out = O.render_chain(wc, wf, g["rays_o"], g["rays_d"], 0., 1., 64, 64, 1024, keep_intermediates=True)
    assert np.array_equal(out["ts"].numpy(), g["ts"])
    assert np.array_equal(out["pts_coarse"].numpy(), g["pts_coarse"])
    np.testing.assert_allclose(out["sigma_coarse"].numpy(), g["sigma_coarse"], atol=2e-4, rtol=1e-5)
    np.testing.assert_allclose(out["rgb_pts_coarse"].numpy(), g["rgb_pts_coarse"], atol=2e-5, rtol=0)
    # teacher-forced resampling on the reference's own coarse weights
    _, ts_f, _, inds, cdf = O.sample_fine(g["rays_o"], g["rays_d"], g["ts"], g["weights_coarse"], 64, return_aux=True)
    assert np.array_equal(cdf.numpy(), g["cdf"])
    assert np.array_equal(inds.numpy(), g["pdf_inds"])
    assert np.array_equal(ts_f.numpy(), g["ts_fine"])
    np.testing.assert_allclose(out["rgb"].numpy(), g["rgb"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(out["depth"].numpy(), g["depth"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(out["acc"].numpy(), g["acc"], atol=1e-4, rtol=0)


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present on this host")
def test_oracle_vs_live_reference():
    """run the real reference chain here and require the oracle to be BIT-identical on every stage."""
    utils, models, dataset, load_llff = ref_import.import_reference()
    wc, wf = weights("w1")
    mc = models.StyleNerf(ref_import.RefArgs, "coarse")
    mf = models.StyleNerf(ref_import.RefArgs, "fine")
    mc.load_state_dict(wc)
    mf.load_state_dict(wf)
    ro, rd = small_rays()
    sel = np.arange(1000, 190000, 1499)[:96]
    o, d = torch.from_numpy(ro[sel]), torch.from_numpy(rd[sel])
    with torch.no_grad():
        pts, ts = utils.sampling_pts_uniform(rays_o=o, rays_d=d, N_samples=64, near=0., far=1.)
        ret = utils.batchify(lambda **kw: mc(**kw), 1024)(pts=pts, dirs=d.unsqueeze(1).expand(96, 64, 3))
        rgb_c, t_c, w_c = utils.alpha_composition(ret["rgb"], ret["sigma"], ts, 0)
        pts_f, ts_f = utils.sampling_pts_fine_torch(o, d, ts, w_c, 64)
        ret_f = utils.batchify(lambda **kw: mf(**kw), 1024)(pts=pts_f, dirs=d.unsqueeze(1).expand(96, 128, 3))
        rgb_f, t_f, w_f = utils.alpha_composition(ret_f["rgb"], ret_f["sigma"], ts_f, 0)
    out = O.render_chain(wc, wf, o, d, 0., 1., 64, 64, 1024, keep_intermediates=True)
    for a, b in ((pts, out["pts_coarse"]), (ts, out["ts"]), (ret["sigma"], out["sigma_coarse"]), (rgb_c, out["rgb_coarse"]),
                 (w_c, out["weights_coarse"]), (ts_f, out["ts_fine"]), (pts_f, out["pts_fine"]), (ret_f["rgb"], out["rgb_pts_fine"]),
                 (rgb_f, out["rgb"]), (t_f, out["depth"]), (w_f, out["weights"])):
        assert torch.equal(a, b)
    # perturb=True replay (utils.py:518-524) and sigma noise replay (utils.py:372-374)
    torch.manual_seed(11)
    p_ref, t_ref = utils.sampling_pts_uniform(rays_o=o, rays_d=d, N_samples=64, near=0., far=1., perturb=True)
    torch.manual_seed(11)
    rand = torch.zeros(96, 64)
    torch.nn.init.uniform_(rand, 0, 1)
    p_o, t_o = O.sample_uniform(o, d, 64, 0., 1., rand=rand)
    assert torch.equal(t_ref, t_o) and torch.equal(p_ref, p_o)
    torch.manual_seed(12)
    ref_noisy = utils.alpha_composition(ret["rgb"], ret["sigma"], ts, 1.0)
    torch.manual_seed(12)
    noise = torch.randn(ret["sigma"].shape) * 1.0
    ora_noisy = O.alpha_composition(ret["rgb"], ret["sigma"], ts, noise=noise)
    assert torch.equal(ref_noisy[0], ora_noisy[0]) and torch.equal(ref_noisy[2], ora_noisy[2])


# ------------------------------------------------------------------ per-ray style head
def test_style_weights_identical_to_reference():
    """oracle.init_style_like_reference(1) == torch.manual_seed(1); StyleMLP_before_concat; StyleMLP_Wild_multilayers."""
    import hashlib
    g = golden("style_chain")
    cs, ws = O.init_style_like_reference(1)
    for tag, sd in (("concat", cs), ("wild", ws)):
        for k, v in sd.items():
            assert hashlib.sha256(np.ascontiguousarray(v.numpy()).tobytes()).hexdigest() == str(g["sha_%s/%s" % (tag, k)]), (tag, k)


def test_style_chain_vs_golden():
    """oracle.render_style_chain against the imported reference's render_style loop body (rendering.py:118-178)."""
    g = golden("style_chain")
    ro, rd = small_rays()
    w0c, w0f = O.init_linear_like_reference(0)
    probe = g["probe_index"]
    wc, wf = O.recalibrate_sigma(w0c, ro[probe], rd[probe]), O.recalibrate_sigma(w0f, ro[probe], rd[probe])
    cs, ws = O.init_style_like_reference(1)
    sel = g["ray_index"]
    out = O.render_style_chain(wc, wf, cs, ws, ro[sel], rd[sel], g["latents"], keep_intermediates=True)
    assert np.array_equal(out["ts_fine"].numpy(), g["ts_fine"])
    for k, tol in (("rgb_coarse", 2e-6), ("weights_coarse", 1e-6), ("rgb", 2e-6), ("depth", 2e-6), ("weights", 1e-6), ("rgb_pts_coarse", 2e-6)):
        np.testing.assert_allclose(out[k].numpy(), g[k], atol=tol, rtol=0, err_msg=k)
    np.testing.assert_allclose(out["concat_features_coarse"][:8].numpy(), g["concat_features_coarse"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(out["rgb_pts_fine"][:16].numpy(), g["rgb_pts_fine"], atol=2e-6, rtol=0)


# ------------------------------------------------------------------ Style_train iteration
def test_style_train_step_vs_golden():
    """oracle.style_train_step_reference against the imported reference's own modules and loss functions wired as
    train_tgtcs.py:354-495 (oracle/make_golden_style_train.py): loss terms, style-module gradients, latent-table gradient."""
    g = golden("style_train")
    ro, rd = small_rays()
    w0c, w0f = weights("w0")
    probe = np.arange(0, ro.shape[0], 743)
    wc = O.recalibrate_sigma(w0c, ro[probe], rd[probe], gain=4.0, shift=10.0)
    wf = O.recalibrate_sigma(w0f, ro[probe], rd[probe], gain=4.0, shift=10.0)
    cs, ws = O.init_style_like_reference(1)
    t = torch.from_numpy

    def batch(tag):
        idx = g[tag + "/ray_index"]
        b = {"rays_o": ro[idx], "rays_d": rd[idx], "rgb_gt": t(g[tag + "/rgb_gt"]), "style_id": t(g[tag + "/style_id"]),
             "frame_id": t(g[tag + "/frame_id"]), "rand": t(g[tag + "/rand"])}
        if tag + "/rgb_origin" in g:
            b["rgb_origin"] = t(g[tag + "/rgb_origin"])
        return b

    prev = (t(g["prev_x"]), t(g["prev_y"]), t(g["prev_x_origin"]))
    losses, gcs, gws, gtab = O.style_train_step_reference(wc, wf, cs, ws, t(g["table"]), t(g["mu"]), t(g["logvar"]), batch("b1"), batch("b2"),
                                                          prev, int(g["frame_num"]), rgb_loss_lambda=1.0, logp_lambda=0.1, loss_coh_lambda=1e2)
    for k in ("loss", "loss_rgb", "loss_logp", "loss_coh"):
        assert abs(losses[k].item() - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), (k, losses[k].item(), float(g[k]))
    np.testing.assert_allclose(gtab.numpy(), g["grad_table"], rtol=2e-4, atol=1e-7)
    for tag, grads in (("concat", gcs), ("wild", gws)):
        for k, gr in grads.items():
            ref_norm = float(g["gnorm_%s/%s" % (tag, k)])
            sl = gr.numpy().reshape(-1)[::97]
            assert abs(np.linalg.norm(gr.numpy().astype(np.float64)) - ref_norm) <= 1e-3 * ref_norm + 1e-9, (tag, k)
            np.testing.assert_allclose(sl, g["gslice_%s/%s" % (tag, k)], rtol=5e-3, atol=1e-5 * ref_norm + 1e-9, err_msg="%s %s" % (tag, k))


def test_philox_known_answers():
    """oracle/philox_oracle.py against the Random123 known-answer vectors of philox4x32-10"""
    import philox_oracle as P
    for ctr, key, exp in P.KAT:
        out = P.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(x) for x in out] == list(exp)
    u = P.uniform(3, 0, 200000)
    z = P.normal(3, 1, 200000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 3e-3
    assert abs(z.mean()) < 1e-2 and abs(z.std() - 1.0) < 1e-2
