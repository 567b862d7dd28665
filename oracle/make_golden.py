"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the REAL reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Every array below is produced by the reference's own functions
(dataset.get_rays_np/ndc_rays_np, utils.sampling_pts_uniform, models.StyleNerf,
utils.batchify, utils.alpha_composition, utils.sampling_pts_fine_torch) driven in
the order of rendering.py:27-51; the searchsorted indices are captured from
inside the reference's own sample_pdf call.  The fixtures travel to the GPU box
(the reference tree does not).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
import render_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

FERN_SMALL = (378, 504, 407.566)   # BASELINE config 1 (H, W, focal)
FERN_FULL = (756, 1008, 815.13)    # BASELINE config 2


def spiral_poses(load_llff, n=120):
    """the synthetic 120-pose path."""
    c2w = np.concatenate([np.eye(4)[:3, :4], np.array([[756.], [1008.], [815.13]])], 1)
    poses = load_llff.render_path_spiral(c2w, np.array([0., 1., 0.]), [0.3, 0.3, 0.05], 3.9, 0.2, .5, 2, n)
    return np.stack(poses, 0)[:, :3, :4]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference_chain(utils, mc, mf, o, d, near=0., far=1., n_c=64, n_f=64, chunk=1024):
    captured = {}
    real_ss = torch.searchsorted

    def spy(*a, **k):
        r = real_ss(*a, **k)
        captured["inds"] = r.clone()
        captured["cdf"] = a[0].clone()
        return r

    n = o.shape[0]
    with torch.no_grad():
        pts, ts = utils.sampling_pts_uniform(rays_o=o, rays_d=d, N_samples=n_c, near=near, far=far)
        fwd = utils.batchify(lambda **kw: mc(**kw), chunk)
        ret = fwd(pts=pts, dirs=d.unsqueeze(1).expand(n, n_c, 3))
        rgb_c, t_c, w_c = utils.alpha_composition(ret["rgb"], ret["sigma"], ts, 0)
        torch.searchsorted = spy
        try:
            pts_f, ts_f = utils.sampling_pts_fine_torch(o, d, ts, w_c, n_f)
        finally:
            torch.searchsorted = real_ss
        fwd_f = utils.batchify(lambda **kw: mf(**kw), chunk)
        ret_f = fwd_f(pts=pts_f, dirs=d.unsqueeze(1).expand(n, n_c + n_f, 3))
        rgb_f, t_f, w_f = utils.alpha_composition(ret_f["rgb"], ret_f["sigma"], ts_f, 0)
    g = dict(rays_o=o, rays_d=d, ts=ts.contiguous(), pts_coarse=pts, rgb_pts_coarse=ret["rgb"], sigma_coarse=ret["sigma"],
             rgb_coarse=rgb_c, depth_coarse=t_c, weights_coarse=w_c, pdf_inds=captured["inds"], cdf=captured["cdf"],
             ts_fine=ts_f, rgb_pts_fine=ret_f["rgb"], sigma_fine=ret_f["sigma"], rgb=rgb_f, depth=t_f, weights=w_f,
             acc=w_f.sum(-1), acc_coarse=w_c.sum(-1))
    return {k: v.numpy() for k, v in g.items()}


def main():
    utils, models, dataset, load_llff = ref_import.import_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- weights: W0 = seed-0 default init in train() order, W1 = sigma-recalibrated
    torch.manual_seed(0)
    mc = models.StyleNerf(ref_import.RefArgs, "coarse")
    mf = models.StyleNerf(ref_import.RefArgs, "fine")
    w0c = {k: v.clone() for k, v in mc.state_dict().items()}
    w0f = {k: v.clone() for k, v in mf.state_dict().items()}

    # ---- rays: reference ray-gen + NDC in fp64, cast fp32 (the contract)
    rays = {}
    meta = {}
    poses = spiral_poses(load_llff)
    for tag, (H, W, f), c2w in (("small_identity", FERN_SMALL, np.eye(4)[:3, :4]),
                                ("full_identity", FERN_FULL, np.eye(4)[:3, :4]),
                                ("full_spiral17", FERN_FULL, poses[17])):
        K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
        ro, rd = dataset.get_rays_np(H, W, K, c2w, False)
        ro, rd = dataset.ndc_rays_np(H, W, K[0][0], 1., ro, rd)
        ro32 = np.ascontiguousarray(ro.reshape(-1, 3), dtype=np.float32)
        rd32 = np.ascontiguousarray(rd.reshape(-1, 3), dtype=np.float32)
        idx = np.unique(np.concatenate([np.arange(0, H * W, 4099), np.arange(0, 2 * W + 3), np.arange(H * W - 17, H * W)]))
        rays[tag + "_idx"] = idx.astype(np.int64)
        rays[tag + "_o"] = ro32[idx]
        rays[tag + "_d"] = rd32[idx]
        rays[tag + "_c2w"] = np.asarray(c2w, np.float64)
        rays[tag + "_hwf"] = np.array([H, W, f], np.float64)
        meta[tag + "_sha_o"] = sha(ro32)
        meta[tag + "_sha_d"] = sha(rd32)
        if tag == "small_identity":
            small = (ro32, rd32)
    rays["spiral_poses"] = poses.astype(np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "rays.npz"), **rays, **{k: np.array(v) for k, v in meta.items()})

    # ---- chain on 160 spread rays of the small frame, W1 (non-degenerate) and W0 (literal default init)
    H, W, f = FERN_SMALL
    ro32, rd32 = small
    sel = np.linspace(0, H * W - 1, 160).astype(np.int64)
    o = torch.from_numpy(ro32[sel])
    d = torch.from_numpy(rd32[sel])
    probe = np.arange(0, H * W, H * W // 300)[:256]
    w1c = O.recalibrate_sigma(w0c, ro32[probe], rd32[probe])
    w1f = O.recalibrate_sigma(w0f, ro32[probe], rd32[probe])
    mc.load_state_dict(w1c)
    mf.load_state_dict(w1f)
    g1 = run_reference_chain(utils, mc, mf, o, d)
    g1["ray_index"] = sel
    g1["probe_index"] = probe
    g1["sha_sigma_w_coarse"] = np.array(sha(w1c["net.sigma_layer.weight"].numpy()))
    g1["sha_sigma_w_fine"] = np.array(sha(w1f["net.sigma_layer.weight"].numpy()))
    np.savez_compressed(os.path.join(GOLDEN, "chain_w1.npz"), **g1)

    mc.load_state_dict(w0c)
    mf.load_state_dict(w0f)
    g0 = run_reference_chain(utils, mc, mf, o[:48], d[:48])
    g0["ray_index"] = sel[:48]
    np.savez_compressed(os.path.join(GOLDEN, "chain_w0.npz"), **g0)

    # ---- weight identity: checksum of every tensor of W0
    wsum = {}
    for tag, sd in (("coarse", w0c), ("fine", w0f)):
        for k, v in sd.items():
            wsum[tag + "/" + k] = np.array(sha(v.numpy()))
    np.savez_compressed(os.path.join(GOLDEN, "weights_w0_sha.npz"), **wsum)

    # ---- stand-alone sample_pdf / compositing vectors with adversarial inputs
    torch.manual_seed(7)
    n = 512
    ts = torch.from_numpy(O.linspace_f32(0., 1., 64)).unsqueeze(0).expand(n, 64).contiguous()
    w = torch.rand(n, 64) ** 8                       # peaky
    w[:32] = 0.                                      # empty rays -> uniform pdf
    w[32:64, 5] = 1.0                                # one-hot
    w[64:96] = w[64:96] * 1e-7                       # tiny weights (denom<1e-5 branch)
    w[96:128, -2] = 50.0                             # mass in the last used bin (u==1 edge)
    torch.searchsorted, real = None, torch.searchsorted
    cap = {}

    def spy(*a, **k):
        r = real(*a, **k)
        cap["inds"] = r.clone()
        return r
    torch.searchsorted = spy
    try:
        _, ts_f = utils.sampling_pts_fine_torch(torch.zeros(n, 3), torch.ones(n, 3), ts, w, 64)
    finally:
        torch.searchsorted = real
    sig = torch.randn(n, 128) * 30.
    rgbp = torch.rand(n, 128, 3)
    rgb_e, t_e, w_e = utils.alpha_composition(rgbp, sig, ts_f, 0)
    rgb_wb, _, _ = utils.alpha_composition(rgbp, sig, ts_f, 0, white_bkgd=True)
    np.savez_compressed(os.path.join(GOLDEN, "stages.npz"), ts=ts.numpy(), weights_in=w.numpy(), pdf_inds=cap["inds"].numpy(),
                        ts_fine=ts_f.numpy(), sigma=sig.numpy(), rgb_pts=rgbp.numpy(), rgb=rgb_e.numpy(), depth=t_e.numpy(),
                        weights=w_e.numpy(), rgb_white=rgb_wb.numpy())
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
