"""TEST INFRASTRUCTURE ONLY (and the CPU baseline of bench.py) -- stages the reference's own hot-path modules for the GPU box.

The reference is pure Python (no build step), so "building" it for the box is a copy: the four modules the ray-render path
lives in (utils.py, models.py, dataset.py, load_llff.py) and the three small modules dataset.py imports are copied, unmodified, from /root/reference -- where they lie --
into oracle/_ref/ (git-ignored build output: it travels with the gpurun snapshot like the built .so, and never enters the
history).  __graft_entry__.build() calls this when /root/reference is present; on the GPU box the prebuilt copy is used as is.
bench.py's `--impl reference` arm and `cpu_baseline` then time the reference's OWN code (kind = "reference"); without the staged
copy they fall back to the oracle port (kind = "port").

    python oracle/stage_ref.py
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# the four hot-path modules + the three small modules dataset.py imports at module scope (dataset.py:8,:16,:17)
# ... and rendering.py: the caller loops (cal_geometry) the drop-in tests drive unchanged
FILES = ("utils.py", "models.py", "dataset.py", "load_llff.py", "VGGNet.py", "Style_function.py", "ray_utils.py", "rendering.py")


def stage(force=False):
    """-> DST when the staged copy exists (made now or earlier), else None."""
    have_src = all(os.path.isfile(os.path.join(SRC, f)) for f in FILES)
    if not have_src:
        return DST if all(os.path.isfile(os.path.join(DST, f)) for f in FILES) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if force or not os.path.isfile(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copyfile(s, d)
        manifest[f] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    return DST


if __name__ == "__main__":
    print(stage(force=True))
