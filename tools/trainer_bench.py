"""NerfTrainer.step timing at several batch sizes (the reference's configs use batch_size 2048).  Development aid."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 32768]
    H, W, f = 756, 1008, 815.13
    wc, wf = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    dev = torch.device("cuda:0")
    r = T.NerfRenderer(dev, mode="bf16")
    ro_all, rd_all = r.raygen(H, W, K, np.eye(4)[:3, :4])
    gen = torch.Generator().manual_seed(4)
    for n in sizes:
        tr = T.NerfTrainer(r, wc, wf, max_rays_per_pass=32768, seed=1)
        batches = []
        for _ in range(4):
            sel = torch.randperm(H * W, generator=gen)[:n].to(dev)
            batches.append((ro_all[sel].contiguous(), rd_all[sel].contiguous(), torch.rand(n, 3, generator=gen).to(dev)))
        for perturb in (False, True):
            for s in range(4):
                tr.step(*batches[s % 4], perturb=perturb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for s in range(20):
                tr.step(*batches[s % 4], perturb=perturb)
            e1.record()
            t_host = (time.perf_counter() - t0) / 20 * 1e3
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("batch %6d rays perturb=%d: %.3f ms per step (host enqueue %.3f ms)  %.0f rays/s" % (n, perturb, ms, t_host, n / ms * 1e3))


if __name__ == "__main__":
    main()
