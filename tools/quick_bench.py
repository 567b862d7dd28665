"""Quick device timing of the individual kernels (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1008 * 64
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16", "fp32"]
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    for mode in modes:
        r = T.NerfRenderer("cuda:0", mode=mode)
        r.set_weights(w0c, w0f)
        ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=0, n=n)
        ms = timeit(lambda: r.nerf_forward_rays(T.NET_COARSE, ro, rd, None, 64, 0., 1.))
        fl = n * 64 * 1186816
        print("%s coarse MLP n=%d: %.3f ms  %.1f TFLOP/s  %.2f Msamples/s" % (mode, n, ms, fl / ms / 1e9, n * 64 / ms / 1e3))
        ts = torch.sort(torch.rand(n, 128, device="cuda"), -1)[0]
        ms = timeit(lambda: r.nerf_forward_rays(T.NET_FINE, ro, rd, ts, 128, 0., 1.))
        fl = n * 128 * 1186816
        print("%s fine MLP n=%d: %.3f ms  %.1f TFLOP/s" % (mode, n, ms, fl / ms / 1e9))
        ms = timeit(lambda: r.render(ro, rd, 0., 1.))
        print("%s render n=%d: %.3f ms  %.3f Mrays/s  %.1f TFLOP/s" % (mode, n, ms, n / ms / 1e3, n * 192 * 1186816 / ms / 1e9))
        r.close()

if __name__ == "__main__":
    main()
