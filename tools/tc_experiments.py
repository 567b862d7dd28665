"""Timing experiments on the tcgen05 MLP kernel with parts of it switched off (results are garbage; timing only)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib
from quick_bench import timeit

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1008 * 128
    flags = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2, 16, 18]
    lib = _lib.load()
    lib.tgtc_debug_tc_flags.argtypes = [ctypes.c_int]
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(w0c, w0f)
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=0, n=n)
    tiles = n * 64 / 128
    for fl in flags:
        lib.tgtc_debug_tc_flags(fl)
        ms = timeit(lambda: r.nerf_forward_rays(T.NET_COARSE, ro, rd, None, 64, 0., 1.), iters=5, warm=2)
        cyc = ms * 1e-3 * 1.9e9 / (tiles / 148)
        print("flags=%2d (%s): %.3f ms  %.1f TFLOP/s-equiv  ~%.0f cycles/tile @1.9GHz" % (
            fl, "+".join(nm for b, nm in ((2, "noEPI"), (16, "noRing")) if fl & b) or "full", ms,
            n * 64 * 1186816 / ms / 1e9, cyc))
    lib.tgtc_debug_tc_flags(0)

if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()
