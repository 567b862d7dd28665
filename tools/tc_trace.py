"""Role timeline of CTA 0 of the tcgen05 MLP kernel (clock64 stamps written by the kernel's trace hook).
Development aid: python tools/tc_trace.py [n_rays] [flags]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1008 * 16
    fl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    lib = _lib.load()
    lib.tgtc_debug_tc_flags.argtypes = [ctypes.c_int]
    lib.tgtc_debug_tc_trace.argtypes = [ctypes.c_void_p]
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(w0c, w0f)
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=0, n=n)
    ts = None if S == 64 else torch.sort(torch.rand(n, S, device="cuda"), -1)[0]
    net = T.NET_COARSE if S == 64 else T.NET_FINE
    for _ in range(2):
        r.nerf_forward_rays(net, ro, rd, ts, S, 0., 1.)
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 4 * 10 * 2 * 2 + 4 * 64, dtype=torch.int64, device="cuda")
    lib.tgtc_debug_tc_flags(fl)
    lib.tgtc_debug_tc_trace(buf.data_ptr())
    r.nerf_forward_rays(net, ro, rd, ts, S, 0., 1.)
    torch.cuda.synchronize()
    lib.tgtc_debug_tc_trace(None)
    lib.tgtc_debug_tc_flags(0)
    full = buf.cpu().numpy()
    tr = full[:640].reshape(4, 4, 10, 2, 2)
    ring = full[640:].reshape(4, 64)
    t0 = tr[tr > 0].min()
    names = ["MMA ", "EPI0", "EPI1", "PE  "]
    ev = []
    for role in range(4):
        for it in range(4):
            for l in range(10):
                for t in range(2):
                    a, b = tr[role, it, l, t]
                    if a > 0:
                        ev.append((a - t0, b - t0, names[role], it, l, t))
    ev.sort()
    print("flags=%d S=%d   begin  end  (dur)   role it layer slot" % (fl, S))
    for a, b, nm, it, l, t in ev:
        print("%8d %8d (%6d)  %s it=%d l=%d t=%d" % (a, b, b - a, nm, it, l, t))
    # summary: per-layer MMA issue->commit duration, epilogue durations, iteration period
    mma = tr[0]
    for it in range(1, 3):
        if mma[it + 1, 0, 0, 0] > 0:
            print("iteration %d period: %d cycles (2 tiles)" % (it, mma[it + 1, 0, 0, 0] - mma[it, 0, 0, 0]))
    d = (tr[1, 1:3, :, :, 1] - tr[1, 1:3, :, :, 0])
    print("EPI0 duration per layer (avg over it=1,2 and slots):", d.mean(axis=(0, 2)).astype(int).tolist())
    gap = tr[1, 1:3, :, :, 0] - tr[0, 1:3, :, :, 1]
    print("AccFull wake - MMA issue end per layer (MMA exec + signal latency):", gap.mean(axis=(0, 2)).astype(int).tolist())
    w = tr[0, 1:3, :, :, 1] - tr[0, 1:3, :, :, 0]
    print("MMA issue duration per layer:", w.mean(axis=(0, 2)).astype(int).tolist())


def ring_report(ring, stages=6):
    if ring.max() == 0:
        return
    r0 = ring[ring > 0].min()
    r = ring - r0
    print("ring path (ns, %globaltimer): chunk | MMA WFull wake | leader prod WEmpty wake | peer prod WEmpty wake | peer fwd WFull wake")
    for j in range(64):
        print("%3d  %7d %7d %7d %7d" % (j, r[0, j], r[1, j], r[2, j], r[3, j]))
    j = np.arange(8, 56)
    print("mean period per chunk (MMA wake to wake): %.1f ns" % np.diff(r[0, 8:56]).mean())
    print("MMA wake(j) -> leader producer WEmpty wake(j+%d): %.1f ns" % (stages, (r[1, j + stages] - r[0, j]).mean()))
    print("MMA wake(j) -> peer   producer WEmpty wake(j+%d): %.1f ns" % (stages, (r[2, j + stages] - r[0, j]).mean()))
    print("peer producer wake(j) -> peer forwarder wake(j): %.1f ns" % (r[3, j] - r[2, j]).mean())
    print("peer forwarder wake(j) -> MMA wake(j): %.1f ns" % (r[0, j] - r[3, j]).mean())
    print("leader producer wake(j) -> MMA wake(j): %.1f ns" % (r[0, j] - r[1, j]).mean())


if __name__ == "__main__":
    main()
