"""A/B on one box: the headline frame (1008x756, 64+128 samples) with compositing fused into the MLP kernel vs the stand-alone
composite kernel (debug hook), alternating, a few rounds.    python tools/ab_fused_composite.py [mode]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib
from bench import synth_nerf_weights, H, W, FOCAL

mode = sys.argv[1] if len(sys.argv) > 1 else "f16"
r = T.NerfRenderer("cuda:0", mode=mode)
r.set_weights(*synth_nerf_weights(0))
K = np.array([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1]])
ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4])
out = r._alloc_out(H * W, 64, 64, False, r.device); out.pop("weights")
lib = _lib.load()
res = {"fused": [], "standalone": []}
for rnd in range(4):
    for name, off in (("fused", 0), ("standalone", 1)):
        lib.tgtc_debug_no_fused_composite(off)
        for _ in range(2):
            r.render(ro, rd, 0., 1., out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.render(ro, rd, 0., 1., out=out)
        e1.record(); torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 5)
lib.tgtc_debug_no_fused_composite(0)
print(json.dumps({"mode": mode, "ms_per_frame": res, "median": {k: float(np.median(v)) for k, v in res.items()}}))
