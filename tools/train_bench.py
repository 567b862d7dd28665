"""Device timing of the training step (forward with stash + backward, both nets).  Development aid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from quick_bench import timeit

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(w0c, w0f)
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4])
    sel = torch.randperm(H * W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))[:n]
    ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
    gt = torch.rand(n, 3, device="cuda")
    grads = torch.zeros(2 * 595844, device="cuda")
    ms = timeit(lambda: r.train_step(ro, rd, gt, grads=grads), iters=5, warm=2)
    fl = n * 192 * 3489024.0
    print("train step n=%d: %.3f ms  %.3f Mrays/s  %.1f TFLOP/s (fwd+dgrad+wgrad algorithmic)" % (n, ms, n / ms / 1e3, fl / ms / 1e9))
    ms_f = timeit(lambda: r.render(ro, rd, 0., 1., want_weights=False), iters=5, warm=2)
    print("inference render n=%d: %.3f ms" % (n, ms_f))

if __name__ == "__main__":
    main()
