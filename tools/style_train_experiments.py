"""Timing experiments on the Style_train forward kernels (results of flagged runs are garbage).  Development aid."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    H, W, f = 756, 1008, 815.13
    wc, wf = B.synth_nerf_weights(0)
    cs, ws = B.synth_style_weights(1)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4])
    sel = torch.randperm(H * W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))[:n]
    ro, rd = ro[sel].contiguous(), rd[sel].contiguous()
    lat = torch.randn(n, 32, device="cuda")
    rand = torch.rand(n, 64, device="cuda")
    gc, gf = torch.randn(n, 3, device="cuda") / n, torch.randn(n, 3, device="cuda") / n
    ws_t = torch.empty(r.style_train_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    lib.tgtc_debug_chain_flags.argtypes = [ctypes.c_int]
    for flags in (0, 1, 4, 8, 13):
        lib.tgtc_debug_chain_flags(flags)
        for _ in range(2):
            fw = r.style_train_forward(ro, rd, lat, rand=rand, workspace=ws_t)
        torch.cuda.synchronize()
        r.profile_enable(True)
        for _ in range(4):
            fw = r.style_train_forward(ro, rd, lat, rand=rand, workspace=ws_t)
            r.style_train_backward(fw["state"], gc, gf)
        torch.cuda.synchronize()
        k = {name: r.profile_read_kind(i) for name, i in (("trunk", 0), ("chain fwd", 1), ("dgrad", 2), ("wgrad+", 3))}
        r.profile_enable(False)
        print("flags %2d:" % flags, {a: round(b[1] / 4, 3) for a, b in k.items()}, "ms per batch (coarse+fine)")
    lib.tgtc_debug_chain_flags(0)


if __name__ == "__main__":
    main()
