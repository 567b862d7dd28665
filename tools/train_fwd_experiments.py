"""Timing experiments on the training forward kernel (results of flagged runs are garbage).  Development aid."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
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
    lib = _lib.load()
    for flags in (0, 32, 64, 96, 128, 224):
        lib.tgtc_debug_tc_flags(flags)
        for _ in range(2):
            r.train_step(ro, rd, gt, grads=grads)
        torch.cuda.synchronize()
        r.profile_enable(True)
        for _ in range(4):
            r.train_step(ro, rd, gt, grads=grads)
        torch.cuda.synchronize()
        k = {name: r.profile_read_kind(i) for name, i in (("fwd", 1), ("dgrad", 2), ("wgrad", 3))}
        r.profile_enable(False)
        print("flags %3d:" % flags, {a: round(b[1] / 4, 3) for a, b in k.items()}, "ms per step (coarse+fine launches)")
    lib.tgtc_debug_tc_flags(0)


if __name__ == "__main__":
    main()
