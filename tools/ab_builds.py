"""A/B of two builds of libtgtc_b200.so on ONE box (box-to-box variation is 1-3 %, larger than most kernel changes): alternates
the headline frame render between the in-tree library and another build, a few rounds each, in separate processes.
    git stash; python tgtc-style_b200/build.py --force; cp tgtc-style_b200/libtgtc_b200.so gpurun_out/lib_base.so; git stash pop; build
    gpurun -- python tools/ab_builds.py gpurun_out/lib_base.so [mode] [workload]
The base library travels in gpurun_out/?  No -- gpurun_out/ is not sent to the box: keep the base build under tools/_ab_base.so
(git-ignored via tools/_ab*, but part of the gpurun snapshot)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
import tgtc_style_b200 as T
from bench import synth_nerf_weights, H, W, FOCAL
mode = sys.argv[1]
r = T.NerfRenderer("cuda:0", mode=mode)
r.set_weights(*synth_nerf_weights(0))
K = np.array([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1]])
ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4])
out = r._alloc_out(H * W, 64, 64, False, r.device); out.pop("weights")
for _ in range(3):
    r.render(ro, rd, 0., 1., out=out)
torch.cuda.synchronize()
ms = []
for rnd in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r.render(ro, rd, 0., 1., out=out)
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / 5)
print(json.dumps(ms))
''' % ROOT


def main():
    base = os.path.abspath(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "f16"
    res = {"new": [], "base": []}
    for rnd in range(3):
        for name, lib in (("new", None), ("base", base)):
            env = dict(os.environ)
            if lib:
                env["TGTC_B200_LIB"] = lib
            else:
                env.pop("TGTC_B200_LIB", None)
            out = subprocess.run([sys.executable, "-c", CHILD, mode], capture_output=True, text=True, env=env, cwd=ROOT)
            if out.returncode != 0:
                print(name, "failed:", out.stderr[-1500:])
                return
            res[name] += json.loads(out.stdout.strip().splitlines()[-1])
    med = {k: sorted(v)[len(v) // 2] for k, v in res.items()}
    print(json.dumps({"mode": mode, "ms_per_frame": res, "median": med, "new_over_base": med["new"] / med["base"]}))


if __name__ == "__main__":
    main()
