"""Training-convergence evidence for Style_train: the same optimisation problem -- fit the two style modules
and a latent table to the stylised renders of a fixed 'teacher' (other style weights, other latents), fixed ray batch with
per-ray (style, frame) ids, no jitter, rgb + logp losses, Adam 5e-4 on the modules and 1e-3 on the table -- run (a) with
StyleTrainer (tcgen05 forward/backward, library loss and latent kernels, fused Adam) on the GPU and (b) with fp32
torch.autograd + torch.optim.Adam through the CPU oracle (the reference's own arithmetic, train_tgtcs.py:404-495).
Prints one JSON document with both loss curves.
    python tools/style_train_convergence.py [steps] [rays]  > profiles/style_train_convergence_r1.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import render_oracle as O
import tgtc_style_b200 as T


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    ro_np, rd_np = O.make_rays(378, 504, 407.566, np.eye(4)[:3, :4])
    probe = np.arange(0, ro_np.shape[0], 743)
    w0c, w0f = O.init_linear_like_reference(0)
    wc = O.recalibrate_sigma(w0c, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0)
    wf = O.recalibrate_sigma(w0f, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0)
    g = torch.Generator().manual_seed(3)
    style_num, frame_num = 2, 4
    sel = np.random.RandomState(1).permutation(ro_np.shape[0])[:n]
    ro, rd = ro_np[sel], rd_np[sel]
    sid, fid = torch.randint(0, style_num, (n,), generator=g), torch.randint(0, frame_num, (n,), generator=g)
    mu, logvar = torch.randn(style_num, 32, generator=g) * 0.3, torch.randn(style_num, 32, generator=g) * 0.2
    # teacher: style weights of seed 5 and its own latent table; student: seed 1 weights, another table
    tcs, tws = O.init_style_like_reference(5)
    t_table = torch.randn(style_num, frame_num, 32, generator=g) * 0.5
    s_table = torch.randn(style_num, frame_num, 32, generator=g) * 0.5
    cs, ws = O.init_style_like_reference(1)
    lat_t = mu[sid] + (t_table.reshape(-1, 32)[sid * frame_num + fid] - mu[sid])
    gt = O.render_style_chain(wc, wf, tcs, tws, ro, rd, lat_t)["rgb"]
    rand = None
    batch = {"rays_o": torch.from_numpy(ro), "rays_d": torch.from_numpy(rd), "rgb_gt": gt, "style_id": sid, "frame_id": fid,
             "rand": torch.full((n, 64), 0.5)}       # mid-bin positions every step: a deterministic problem for both runs
    # (a) GPU
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(wc, wf)
    dev = r.device
    lat = T.StyleLatents(s_table.to(dev), mu.to(dev), logvar.to(dev), dataset_type="blender")
    tr = T.StyleTrainer(r, cs, ws, lat, lr=5e-4, frame_num=frame_num)
    b_dev = {k: v.to(dev) for k, v in batch.items()}
    t0 = time.time()
    gpu_curve, gpu_rgb = [], []
    for _ in range(steps):
        out = tr.step(b_dev)
        gpu_curve.append(out["loss"].item())
        gpu_rgb.append(out["loss_rgb"].item())
    gpu_s = time.time() - t0
    # (b) CPU oracle: autograd + torch Adam on fp32 masters
    pc = {k: torch.nn.Parameter(v.clone()) for k, v in cs.items()}
    pw = {k: torch.nn.Parameter(v.clone()) for k, v in ws.items()}
    tab = torch.nn.Parameter(s_table.clone())
    opt = torch.optim.Adam(list(pc.values()) + list(pw.values()), lr=5e-4, betas=(0.9, 0.999))
    opt_l = torch.optim.Adam([tab], lr=1e-3)
    cpu_curve, cpu_rgb = [], []
    t0 = time.time()
    for step in range(steps):
        losses, gc, gw, gtab = O.style_train_step_reference(wc, wf, {k: v.detach() for k, v in pc.items()}, {k: v.detach() for k, v in pw.items()},
                                                            tab.detach(), mu, logvar, batch, None, None, frame_num, dataset_type="blender")
        for k in pc:
            pc[k].grad = gc[k]
        for k in pw:
            pw[k].grad = gw[k]
        tab.grad = gtab
        opt.step()
        opt_l.step()
        cpu_curve.append(losses["loss"].item())
        cpu_rgb.append(losses["loss_rgb"].item())
    cpu_s = time.time() - t0
    rel = [abs(a - b) / max(b, 1e-12) for a, b in zip(gpu_curve, cpu_curve)]
    rel_rgb = [abs(a - b) / max(b, 1e-12) for a, b in zip(gpu_rgb, cpu_rgb)]
    print(json.dumps({"problem": "fit both style modules + a %dx%d latent table to a teacher's stylised render, %d fixed rays with per-ray "
                                 "(style, frame) ids, rgb + logp losses, Adam 5e-4 / 1e-3" % (style_num, frame_num, n), "steps": steps,
                      "loss_bf16_tcgen05_gpu": gpu_curve, "loss_fp32_autograd_cpu_oracle": cpu_curve,
                      "loss_rgb_bf16_tcgen05_gpu": gpu_rgb, "loss_rgb_fp32_autograd_cpu_oracle": cpu_rgb,
                      "max_rel_loss_diff": max(rel), "final_rel_loss_diff": rel[-1], "max_rel_loss_rgb_diff": max(rel_rgb),
                      "loss_rgb_reduction_gpu": gpu_rgb[-1] / gpu_rgb[0], "loss_rgb_reduction_cpu": cpu_rgb[-1] / cpu_rgb[0],
                      "loss_reduction_gpu": gpu_curve[-1] / gpu_curve[0], "loss_reduction_cpu": cpu_curve[-1] / cpu_curve[0],
                      "seconds_gpu": gpu_s, "seconds_cpu": cpu_s}, indent=1))


if __name__ == "__main__":
    main()
