"""Multi-GPU correctness check, run under torchrun on N GPUs of one box:
  * a frame rendered as N ray shards + NCCL tile all-gather == the same frame rendered whole on rank 0's GPU, bit for bit
  * a training step with the batch split over N ranks + ONE gradient all-reduce == the single-GPU step (fp32 reduction
    tolerance), and every rank ends with identical parameters
  * two Style_train iterations (the second with the coherence term) with both batches split over N ranks == the single-GPU
    iterations: losses, style-module gradients, latent table
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import render_oracle as O
import tgtc_style_b200 as T


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, f = 378, 504, 407.566
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    w0c, w0f = O.init_linear_like_reference(0)
    ro_np, rd_np = O.make_rays(H, W, f, np.eye(4)[:3, :4])
    probe = np.arange(0, H * W, 743)
    wc, wf = O.recalibrate_sigma(w0c, ro_np[probe], rd_np[probe]), O.recalibrate_sigma(w0f, ro_np[probe], rd_np[probe])
    r = T.NerfRenderer(device=dev, mode="bf16")
    r.set_weights(wc, wf)
    res = {"world": world}
    # ---- render: shards + all-gather vs whole frame
    pose = np.array([[1, 0, 0, 0.05], [0, 1, 0, -0.02], [0, 0, 1, 0.01]], dtype=np.float64)
    n = H * W
    frames = dict(T.render_path_sharded(r, H, W, K, [pose], split="rows"))
    whole = r.render_frame(H, W, K, pose)
    res["render_bit_identical"] = all(bool(torch.equal(frames[0][k], whole[k])) for k in ("rgb", "depth", "acc"))
    fr2 = dict(T.render_path_sharded(r, H, W, K, [pose, np.eye(4)[:3, :4], pose], split="frames"))
    res["frames_split_bit_identical"] = bool(torch.equal(fr2[0]["rgb"], whole["rgb"])) and bool(torch.equal(fr2[2]["rgb"], whole["rgb"])) and len(fr2) == 3
    # ---- training: N-way split + all-reduce vs single GPU
    nb = 2048
    g = torch.Generator().manual_seed(4)
    sel = torch.randperm(n, generator=g)[:nb]
    ro, rd = torch.from_numpy(ro_np)[sel].to(dev), torch.from_numpy(rd_np)[sel].to(dev)
    gt = torch.rand(nb, 3, generator=g).to(dev)
    tr = T.NerfTrainer(r, wc, wf)
    tr.step(ro, rd, gt)                       # every rank passes the global batch and takes its shard
    g_multi = tr.grads.clone()
    p_multi = torch.cat([p.detach().flatten() for d in tr.params for p in d.values()])
    r1 = T.NerfRenderer(device=dev, mode="bf16")
    single = T.NerfTrainer(r1, wc, wf, group=dist.new_group([rank]))    # a 1-rank group: no exchange
    single.step(ro, rd, gt)
    g_single = single.grads
    p_single = torch.cat([p.detach().flatten() for d in single.params for p in d.values()])
    res["grad_rel_err_vs_single_gpu"] = float((g_multi - g_single).norm() / g_single.norm())
    res["param_max_abs_diff_vs_single_gpu"] = float((p_multi - p_single).abs().max())
    pm = [torch.empty_like(p_multi) for _ in range(world)]
    dist.all_gather(pm, p_multi)
    res["params_identical_across_ranks"] = all(bool(torch.equal(pm[0], x)) for x in pm)
    # ---- Style_train: both batches of an iteration split over the ranks vs one GPU
    cs, ws = O.init_style_like_reference(1)
    ns = 512
    table = torch.randn(2, 5, 32, generator=g) * 0.5
    mu, logvar = torch.randn(2, 32, generator=g) * 0.3, torch.randn(2, 32, generator=g) * 0.2
    its = []
    for it in range(2):
        pair = []
        for origin in (False, True):
            idx = torch.randperm(n, generator=g)[:ns]
            b = {"rays_o": torch.from_numpy(ro_np)[idx].to(dev), "rays_d": torch.from_numpy(rd_np)[idx].to(dev),
                 "rgb_gt": torch.rand(ns, 3, generator=g).to(dev), "style_id": torch.randint(0, 2, (ns,), generator=g).to(dev),
                 "frame_id": torch.randint(0, 5, (ns,), generator=g).to(dev), "rand": torch.rand(ns, 64, generator=g).to(dev)}
            if origin:
                b["rgb_origin"] = torch.rand(ns, 3, generator=g).to(dev)
            pair.append(b)
        its.append(pair)

    def style_run(renderer, group, lo, hi):
        lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev))
        st = T.StyleTrainer(renderer, cs, ws, lat, frame_num=5, group=group)
        out = None
        for pair in its:
            out = st.step(*({k: v[lo:hi] for k, v in d.items()} for d in pair))
        torch.cuda.synchronize()
        return st.grads.clone(), lat.latents.detach().clone(), {k: float(v) for k, v in out.items()}

    r.set_weights(wc, wf)
    b0, b1 = rank * ns // world, (rank + 1) * ns // world
    gm, tm, lm = style_run(r, None, b0, b1)
    r1.set_weights(wc, wf)
    gs, ts_, ls = style_run(r1, dist.new_group([rank]), 0, ns)
    res["style_grad_rel_err_vs_single_gpu"] = float((gm - gs).norm() / gs.norm())
    res["style_table_max_abs_diff_vs_single_gpu"] = float((tm - ts_).abs().max())
    res["style_losses"] = {"multi": lm, "single": ls}
    style_ok = (res["style_grad_rel_err_vs_single_gpu"] < 1e-3 and res["style_table_max_abs_diff_vs_single_gpu"] < 2e-3
                and ls["loss_coh"] > 0 and abs(lm["loss"] - ls["loss"]) <= 1e-3 * max(1.0, abs(ls["loss"])))
    ok = (res["render_bit_identical"] and res["frames_split_bit_identical"] and res["grad_rel_err_vs_single_gpu"] < 1e-4
          and res["params_identical_across_ranks"] and style_ok)
    res["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
