"""Training-convergence evidence for the bf16 training step: the same optimisation problem -- fit the
coarse+fine NeRF pair to the renders of a fixed 'teacher' pair, fixed ray batch, Adam lr 5e-4 -- run (a) with NerfTrainer
(tcgen05 forward/backward, fused Adam) on the GPU and (b) with fp32 torch.autograd + torch.optim.Adam through the CPU oracle
(the reference's own training arithmetic, train_tgtcs.py:228-276).  Prints one JSON document with both loss curves.
    python tools/train_convergence.py [steps] [rays]  > profiles/train_convergence_r1.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import render_oracle as O
import tgtc_style_b200 as T


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ro_np, rd_np = O.make_rays(378, 504, 407.566, np.eye(4)[:3, :4])
    probe = np.arange(0, ro_np.shape[0], 743)
    w0c, w0f = O.init_linear_like_reference(0)
    # teacher: a smooth non-degenerate field (sigma ~ N(10, 4^2)); student: the same architecture from another seed
    tc, tf = O.recalibrate_sigma(w0c, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0), O.recalibrate_sigma(w0f, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0)
    s1c, s1f = O.init_linear_like_reference(7)
    sc, sf = O.recalibrate_sigma(s1c, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0), O.recalibrate_sigma(s1f, ro_np[probe], rd_np[probe], gain=4.0, shift=10.0)
    sel = np.random.RandomState(1).permutation(ro_np.shape[0])[:n]
    ro, rd = ro_np[sel], rd_np[sel]
    gt = O.render_chain(tc, tf, ro, rd, 0., 1., 64, 64, 1024)["rgb"]
    # (a) GPU
    r = T.NerfRenderer("cuda:0", mode="bf16")
    tr = T.NerfTrainer(r, sc, sf, lr=5e-4)
    ro_d, rd_d, gt_d = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda(), gt.cuda()
    t0 = time.time()
    gpu_curve = [tr.step(ro_d, rd_d, gt_d).item() for _ in range(steps)]
    gpu_s = time.time() - t0
    # (b) CPU oracle: autograd + torch Adam on fp32 masters
    pc = {k: torch.nn.Parameter(v.clone()) for k, v in sc.items()}
    pf = {k: torch.nn.Parameter(v.clone()) for k, v in sf.items()}
    opt = torch.optim.Adam(list(pc.values()) + list(pf.values()), lr=5e-4, betas=(0.9, 0.999))
    cpu_curve = []
    t0 = time.time()
    for step in range(steps):
        loss, gc, gf, _, _, _ = O.train_step_reference({k: v.detach() for k, v in pc.items()}, {k: v.detach() for k, v in pf.items()}, ro, rd, gt)
        for k in pc:
            pc[k].grad = gc[k]
        for k in pf:
            pf[k].grad = gf[k]
        opt.step()
        lr = 5e-4 * (0.1 ** ((step + 1) / 100000))
        for g in opt.param_groups:
            g["lr"] = lr
        cpu_curve.append(loss.item())
    cpu_s = time.time() - t0
    rel = [abs(a - b) / max(b, 1e-12) for a, b in zip(gpu_curve, cpu_curve)]
    print(json.dumps({"problem": "fit coarse+fine NeRF to a teacher pair's render, %d fixed rays, Adam lr 5e-4, perturb=0" % n, "steps": steps,
                      "loss_bf16_tcgen05_gpu": gpu_curve, "loss_fp32_autograd_cpu_oracle": cpu_curve,
                      "max_rel_loss_diff": max(rel), "final_rel_loss_diff": rel[-1],
                      "loss_reduction_gpu": gpu_curve[-1] / gpu_curve[0], "loss_reduction_cpu": cpu_curve[-1] / cpu_curve[0],
                      "seconds_gpu": gpu_s, "seconds_cpu": cpu_s}, indent=1))


if __name__ == "__main__":
    main()
