"""TEST INFRASTRUCTURE ONLY -- golden vectors of the per-ray style head, produced by the IMPORTED reference
(models.StyleMLP_before_concat, models.StyleMLP_Wild_multilayers, called as in rendering.py:118-178 with perturb=False).

    python oracle/make_golden_style.py        # in the build container (needs /root/reference)

Writes tests/golden/style_chain.npz.  tests/test_oracle_golden.py checks oracle.render_style_chain against it everywhere.
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


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@torch.no_grad()
def main():
    utils, models, dataset, load_llff = ref_import.import_reference()
    A = ref_import.RefArgs
    torch.manual_seed(0)
    mc = models.StyleNerf(A, "coarse")
    mf = models.StyleNerf(A, "fine")
    torch.manual_seed(1)
    cs = models.StyleMLP_before_concat(A)
    ws = models.StyleMLP_Wild_multilayers(A)
    H, W, f = 378, 504, 407.566
    ro, rd = O.make_rays(H, W, f, np.eye(4)[:3, :4])
    probe = np.arange(0, H * W, H * W // 300)[:256]
    w0c, w0f = {k: v.clone() for k, v in mc.state_dict().items()}, {k: v.clone() for k, v in mf.state_dict().items()}
    mc.load_state_dict(O.recalibrate_sigma(w0c, ro[probe], rd[probe]))
    mf.load_state_dict(O.recalibrate_sigma(w0f, ro[probe], rd[probe]))
    sel = np.linspace(0, H * W - 1, 96).astype(np.int64)
    o, d = torch.from_numpy(ro[sel]), torch.from_numpy(rd[sel])
    n = o.shape[0]
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(1, 32, generator=g).expand(n, 32).contiguous()      # one (style, frame) for the batch
    lat2 = torch.mean(lat, dim=1, keepdims=True)                          # rendering.py:126

    def one_pass(model, pts, S):
        ret = model(pts=pts, dirs=d.unsqueeze(1).expand(n, S, 3))
        l1 = lat.unsqueeze(1).expand(n, S, 32)                             # rendering.py:127
        cf = cs(x=ret["pts"], latent=l1)["concat_features"]               # rendering.py:129-130
        concated = torch.concat((ret["base_remap"], cf), dim=-1)          # rendering.py:132
        l2 = torch.unsqueeze(lat2, dim=2).expand(n, S, 32)                # rendering.py:139
        rgb = ws(x=ret["pts"], concated=concated, latent=l2)["rgb"]       # rendering.py:140
        return ret, cf, rgb

    pts, ts = utils.sampling_pts_uniform(rays_o=o, rays_d=d, N_samples=64, near=0., far=1., perturb=False)
    ret, cf, rgb_s = one_pass(mc, pts, 64)
    rgb_c, t_c, w_c = utils.alpha_composition(rgb_s, ret["sigma"], ts, 0)
    pts_f, ts_f = utils.sampling_pts_fine_torch(o, d, ts, w_c, 64)
    ret_f, cf_f, rgb_sf = one_pass(mf, pts_f, 128)
    rgb_f, t_f, w_f = utils.alpha_composition(rgb_sf, ret_f["sigma"], ts_f, 0)
    out = dict(ray_index=sel, probe_index=probe, latents=lat.numpy(), rgb_coarse=rgb_c.numpy(), weights_coarse=w_c.numpy(),
               concat_features_coarse=cf[:8].numpy(), rgb_pts_coarse=rgb_s.numpy(), ts_fine=ts_f.numpy(), rgb=rgb_f.numpy(),
               depth=t_f.numpy(), weights=w_f.numpy(), rgb_pts_fine=rgb_sf[:16].numpy())
    for tag, m in (("concat", cs), ("wild", ws)):
        for k, v in m.state_dict().items():
            out["sha_%s/%s" % (tag, k)] = np.array(sha(v.numpy()))
    np.savez_compressed(os.path.join(GOLDEN, "style_chain.npz"), **out)
    print("style_chain.npz", os.path.getsize(os.path.join(GOLDEN, "style_chain.npz")))


if __name__ == "__main__":
    main()
