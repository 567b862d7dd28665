"""TEST INFRASTRUCTURE ONLY -- imports the real reference from /root/reference.

Only usable inside the build container (the GPU box has no /root/reference).
Used by oracle/make_golden.py to produce tests/golden/*.npz and by the CPU test
that pins oracle/render_oracle.py against the reference's own arithmetic.

The reference's utils.py imports I/O-only packages that are not installed
(imageio, plyfile, pyrender, matplotlib, skimage, natsort: utils.py:7-19);
they are stubbed with empty modules.  Nothing on the
ray-rendering path touches them.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns (utils, models, dataset, load_llff) modules of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _stub("imageio")
    _stub("plyfile", PlyElement=object, PlyData=object)
    _stub("pyrender")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = plt
    sk = _stub("skimage")
    skf = _stub("skimage.feature")
    if not hasattr(sk, "feature"):
        sk.feature = skf
    if not hasattr(skf, "canny"):
        skf.canny = None
    _stub("natsort", natsorted=sorted)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's module names are generic ("utils", "models", ...): make
    # sure we get ITS files and not something else already imported
    mods = []
    for name in ("utils", "models", "dataset", "load_llff"):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
        mods.append(importlib.import_module(name))
    return tuple(mods)


class RefArgs:
    """The hot-path flags of configs/fern.txt + config.py defaults."""
    use_viewdir = True
    act_type = "relu"
    embed_freq_coor = 10
    embed_freq_dir = 4
    netdepth = 8
    netdepth_fine = 8
    netwidth = 256
    netwidth_fine = 256
    siren_sigma_mul = 0.0
    style_D = 8          # configs/fern.txt:26
    vae_latent = 32      # config.py:86
