"""Role timeline of CTA 0 of the dgrad kernel (development aid)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    lib = _lib.load()
    lib.tgtc_debug_bwd_trace.argtypes = [ctypes.c_void_p]
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(w0c, w0f)
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=0, n=n)
    gt = torch.rand(n, 3, device="cuda")
    for _ in range(2):
        r.train_step(ro, rd, gt)
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 4 * 9 * 2 * 3, dtype=torch.int64, device="cuda")
    lib.tgtc_debug_bwd_trace(buf.data_ptr())
    r.train_step(ro, rd, gt)
    torch.cuda.synchronize()
    lib.tgtc_debug_bwd_trace(None)
    tr = buf.cpu().numpy().reshape(4, 4, 9, 2, 3)   # the fine pass overwrote the coarse pass: this is the fine net's dgrad
    t0 = tr[tr > 0].min()
    ev = []
    for role, nm in ((0, "MMA"), (1, "EPI"), (2, "IN ")):
        for it in range(4):
            for g in range(9):
                for t in range(2):
                    a = tr[role, it, g, t]
                    if a[0] > 0:
                        ev.append((a[0] - t0, nm, it, g, t, [int(x - t0) if x > 0 else -1 for x in a]))
    ev.sort()
    for e in ev[:110]:
        print("%8d %s it=%d g=%d t=%d  %s" % (e[0], e[1], e[2], e[3], e[4], e[5]))
    m = tr[0]
    for it in range(1, 3):
        if m[it + 1, 0, 0, 0] > 0:
            print("iteration %d period: %d cycles" % (it, m[it + 1, 0, 0, 0] - m[it, 0, 0, 0]))
    e = tr[1, 1:3]
    print("EPI: AccFull wake -> mask tile landed:", (e[..., 1] - e[..., 0]).mean(axis=(0, 2)).astype(int).tolist())
    print("EPI: mask landed -> done:", (e[..., 2] - e[..., 1]).mean(axis=(0, 2)).astype(int).tolist())
    mm = tr[0, 1:3]
    print("MMA issue duration:", (mm[..., 1] - mm[..., 0]).mean(axis=(0, 2)).astype(int).tolist())

if __name__ == "__main__":
    main()
