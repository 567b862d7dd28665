"""TEST INFRASTRUCTURE ONLY -- golden vectors of one Style_train iteration (train_tgtcs.py:354-495), produced by the IMPORTED
reference: models.StyleNerf x 2 (frozen), models.StyleMLP_before_concat, models.StyleMLP_Wild_multilayers,
models.StyleLatents_variational (forward + minus_logp), utils.sampling_pts_uniform(perturb=True), utils.alpha_composition,
utils.sampling_pts_fine_torch, utils.img2mse, utils.L2_norm, VGGNet.cosine_similarity -- wired as the loop body wires them,
with the previous loss_coh batch's maps as constants (the loop as written backpropagates through stale graphs and does not
run on torch >= 1.5; DESIGN.md "Style_train").

    python oracle/make_golden_style_train.py        # in the build container (needs /root/reference)

Writes tests/golden/style_train.npz: inputs, the loss terms, gradient checksums/slices for both style modules and the full
gradient of the latent table.  tests/test_oracle_golden.py checks oracle.style_train_step_reference against it.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402
import render_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    utils, models, dataset, load_llff = ref_import.import_reference()
    import VGGNet
    A = ref_import.RefArgs
    torch.manual_seed(0)
    mc = models.StyleNerf(A, "coarse")
    mf = models.StyleNerf(A, "fine")
    torch.manual_seed(1)
    cs = models.StyleMLP_before_concat(A)
    ws = models.StyleMLP_Wild_multilayers(A)
    H, W, f = 378, 504, 407.566
    ro, rd = O.make_rays(H, W, f, np.eye(4)[:3, :4])
    probe = np.arange(0, H * W, 743)
    mc.load_state_dict(O.recalibrate_sigma({k: v.clone() for k, v in mc.state_dict().items()}, ro[probe], rd[probe], gain=4.0, shift=10.0))
    mf.load_state_dict(O.recalibrate_sigma({k: v.clone() for k, v in mf.state_dict().items()}, ro[probe], rd[probe], gain=4.0, shift=10.0))
    style_num, frame_num, n = 2, 5, 24
    g = torch.Generator().manual_seed(33)
    lm = models.StyleLatents_variational(style_num=style_num, frame_num=frame_num, latent_dim=32)
    with torch.no_grad():
        lm.latents.copy_(torch.randn(style_num, frame_num, 32, generator=g) * 0.5)
        lm.style_latents_mu.copy_(torch.randn(style_num, 32, generator=g) * 0.3)
        lm.style_latents_logvar.copy_(torch.randn(style_num, 32, generator=g) * 0.2)

    def batch(origin):
        idx = torch.randperm(H * W, generator=g)[:n].numpy()
        b = {"ray_index": idx, "rgb_gt": torch.rand(n, 3, generator=g), "style_id": torch.randint(0, style_num, (n,), generator=g),
             "frame_id": torch.randint(0, frame_num, (n,), generator=g), "seed": int(torch.randint(0, 10000, (1,), generator=g))}
        if origin:
            b["rgb_origin"] = torch.rand(n, 3, generator=g)
        return b

    b1, b2 = batch(False), batch(True)
    prev = tuple(torch.rand(n, 3, generator=g) for _ in range(3))        # x, y, x_origin of the previous loss_coh batch

    def forward(b):
        """train_tgtcs.py:362-381 / :404-424 and :431-479 for one batch, the reference's own functions"""
        o, d = torch.from_numpy(ro[b["ray_index"]]), torch.from_numpy(rd[b["ray_index"]])
        torch.manual_seed(b["seed"])
        pts, ts = utils.sampling_pts_uniform(rays_o=o, rays_d=d, N_samples=64, near=0., far=1., perturb=True)
        torch.manual_seed(b["seed"])
        rand = torch.zeros([n, 64])
        torch.nn.init.uniform_(rand, 0, 1)                                # the uniforms the call above drew (utils.py:519-520)
        b["rand"] = rand
        first = lm(style_ids=b["style_id"], frame_ids=b["frame_id"], type="llff")
        lat2 = torch.mean(first, dim=1, keepdims=True)

        def one_pass(model, pts, S):
            with torch.no_grad():
                ret = model(pts=pts, dirs=d.unsqueeze(1).expand(n, S, 3))
            cf = cs(x=ret["pts"], latent=first.unsqueeze(1).expand(n, S, 32))["concat_features"]
            concated = torch.cat((ret["base_remap"], cf), dim=-1)
            return ret, ws(x=ret["pts"], concated=concated, latent=torch.unsqueeze(lat2, dim=2).expand(n, S, 32))["rgb"]

        ret, rgb_s = one_pass(mc, pts, 64)
        rgb_c, _, w_c = utils.alpha_composition(rgb_s, ret["sigma"], ts, 0)
        pts_f, ts_f = utils.sampling_pts_fine_torch(o, d, ts, w_c, 64)
        ret_f, rgb_sf = one_pass(mf, pts_f, 128)
        rgb_f, _, _ = utils.alpha_composition(rgb_sf, ret_f["sigma"], ts_f, 0)
        return rgb_c, rgb_f

    c2, f2 = forward(b2)
    x, y, x_org = prev
    loss_coh = utils.L2_norm(VGGNet.cosine_similarity(c2, x) - VGGNet.cosine_similarity(b2["rgb_origin"], x_org))          # :401
    x_org_now = b2["rgb_origin"]                                                                                              # :403
    loss_coh = loss_coh + utils.L2_norm(VGGNet.cosine_similarity(f2, y) - VGGNet.cosine_similarity(b2["rgb_origin"], x_org_now))  # :456
    rgb_c, rgb_f = forward(b1)
    loss_rgb = 1.0 * utils.img2mse(rgb_c, b1["rgb_gt"]) + 1.0 * utils.img2mse(rgb_f, b1["rgb_gt"])                          # :425, :480-481
    loss_logp = 0.1 * lm.minus_logp(style_ids=b1["style_id"], frame_ids=b1["frame_id"], data_type="llff")                   # :426-427
    loss = loss_rgb + loss_logp + 1e2 * loss_coh                                                                              # loss_for_style, :482
    loss.backward(retain_graph=True)                     # style_optimizer: loss_for_style.backward (:490) -> the two style modules
    # the latent table: latents_model_1.optimize(loss) with loss = loss_rgb + loss_logp (:481, :495; models.py:544-549) zeroes
    # the table's gradient and backpropagates WITHOUT the coherence term
    grad_table = torch.autograd.grad(loss_rgb + loss_logp, lm.latents)[0]
    out = dict(style_num=style_num, frame_num=frame_num, table=lm.latents.detach().numpy(), mu=lm.style_latents_mu.detach().numpy(),
               logvar=lm.style_latents_logvar.detach().numpy(), prev_x=x.numpy(), prev_y=y.numpy(), prev_x_origin=x_org.numpy(),
               loss=loss.item(), loss_rgb=loss_rgb.item(), loss_logp=loss_logp.item(), loss_coh=loss_coh.item(),
               rgb_coarse=rgb_c.detach().numpy(), rgb_fine=rgb_f.detach().numpy(), coh_rgb_coarse=c2.detach().numpy(),
               coh_rgb_fine=f2.detach().numpy(), grad_table=grad_table.numpy())
    for tag, b in (("b1", b1), ("b2", b2)):
        for k, v in b.items():
            if k != "seed":
                out["%s/%s" % (tag, k)] = v.numpy() if torch.is_tensor(v) else np.asarray(v)
    for tag, m in (("concat", cs), ("wild", ws)):
        for k, p in m.named_parameters():
            gr = p.grad.numpy()
            out["gnorm_%s/%s" % (tag, k)] = np.array(np.linalg.norm(gr.astype(np.float64)))
            out["gslice_%s/%s" % (tag, k)] = gr.reshape(-1)[::97].copy()      # every 97th element
    np.savez_compressed(os.path.join(GOLDEN, "style_train.npz"), **out)
    print("style_train.npz", os.path.getsize(os.path.join(GOLDEN, "style_train.npz")), "loss", loss.item(), loss_rgb.item(), loss_logp.item(),
          loss_coh.item())


if __name__ == "__main__":
    main()
