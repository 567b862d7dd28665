"""StyleTrainer.step timing at several batch sizes (the reference's configs use batch_size_style 256 / 1024).  Development aid."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]
    H, W, f = 756, 1008, 815.13
    wc, wf = B.synth_nerf_weights(0)
    cs, ws = B.synth_style_weights(1)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    dev = torch.device("cuda:0")
    r = T.NerfRenderer(dev, mode="bf16")
    r.set_weights(wc, wf)
    ro_all, rd_all = r.raygen(H, W, K, np.eye(4)[:3, :4])
    gen = torch.Generator().manual_seed(4)
    table = torch.randn(4, 20, 32, generator=gen) * 0.5
    mu, logvar = torch.randn(4, 32, generator=gen) * 0.3, torch.randn(4, 32, generator=gen) * 0.2
    for n in sizes:
        lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev))
        tr = T.StyleTrainer(r, cs, ws, lat, frame_num=20)
        batches = B._style_train_batches(ro_all, rd_all, n, 4, 20, gen, dev)
        for s in range(4):
            tr.step(*batches[s % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for s in range(20):
            tr.step(*batches[s % 4])
        e1.record()
        t_host = (time.perf_counter() - t0) / 20 * 1e3
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("batch %5d rays x2: %.3f ms per iteration (host enqueue %.3f ms)  %.0f rays/s" % (n, ms, t_host, 2 * n / ms * 1e3))


if __name__ == "__main__":
    main()
