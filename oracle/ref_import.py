"""TEST INFRASTRUCTURE ONLY -- imports the real reference: from /root/reference in the build
container, else from the staged copy oracle/_ref/ (oracle/stage_ref.py; git-ignored, travels to
the GPU box with the snapshot).
Used by oracle/make_golden.py to produce tests/golden/*.npz, by the CPU test that pins
oracle/render_oracle.py against the reference's own arithmetic, and by bench.py's reference arm /
cpu_baseline leg (kind = "reference").

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

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if os.path.isfile("/root/reference/utils.py") else _STAGED


def reference_available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in ("utils.py", "models.py", "dataset.py", "load_llff.py"))


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


def import_rendering():
    """The reference's rendering module (cal_geometry, render_style, render_train_style: the caller loops of the hot path);
    its only imports are `from utils import *` and time."""
    import_reference()
    m = sys.modules.get("rendering")
    if m is not None and not getattr(m, "__file__", "").startswith(REFERENCE_ROOT):
        del sys.modules["rendering"]
    return importlib.import_module("rendering")


def import_reference():
    """Returns (utils, models, dataset, load_llff) modules of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _stub("imageio", imwrite=lambda *a, **k: None, mimwrite=lambda *a, **k: None)
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


def reference_chain(utils, model, model_fine, rays_o, rays_d, near=0., far=1., n_samples=64, n_fine=64, chunk=1024):
    """The loop body of rendering.py:27-51 (cal_geometry) for one batch of rays, driving the reference's OWN callables exactly
    as train_tgtcs.py:14-16,:30,:37 binds them.  -> (rgb [N,3], depth [N], weights [N,S+F]) torch tensors."""
    import torch
    n = rays_o.shape[0]
    with torch.no_grad():
        pts, ts = utils.sampling_pts_uniform(rays_o=rays_o, rays_d=rays_d, N_samples=n_samples, near=near, far=far)
        ret = utils.batchify(lambda **kw: model(**kw), chunk)(pts=pts, dirs=rays_d.unsqueeze(1).expand(n, n_samples, 3))
        _, _, w_c = utils.alpha_composition(ret["rgb"], ret["sigma"], ts, 0)
        pts_f, ts_f = utils.sampling_pts_fine_torch(rays_o, rays_d, ts, w_c, n_fine)
        ret_f = utils.batchify(lambda **kw: model_fine(**kw), chunk)(pts=pts_f, dirs=rays_d.unsqueeze(1).expand(n, n_samples + n_fine, 3))
        return utils.alpha_composition(ret_f["rgb"], ret_f["sigma"], ts_f, 0)


def reference_nets(models, seed=0):
    """torch.manual_seed(seed); StyleNerf('coarse'); StyleNerf('fine') -- train_tgtcs.py:25-37 order (weight set W0 for seed 0)."""
    import torch
    torch.manual_seed(seed)
    return models.StyleNerf(RefArgs, "coarse"), models.StyleNerf(RefArgs, "fine")


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
