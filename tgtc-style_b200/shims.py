"""Drop-in callables with the reference's own signatures.

The reference injects four callables into its render/training loops:
  samp_func        = utils.sampling_pts_uniform        (train_tgtcs.py:14)
  samp_func_fine   = utils.sampling_pts_fine_torch     (train_tgtcs.py:16)
  model_forward    = utils.batchify(lambda **kw: model(**kw), chunk)   (train_tgtcs.py:30, :37)
  alpha_composition (module global via `from utils import *`, rendering.py:1)
`make_callables(renderer)` returns B200 versions with the same names, argument
meaning and return structure; `patch(renderer, modules)` rebinds them in the
reference's module globals (star-imports copy names, so each module's copy is
patched) so train_tgtcs.py / rendering.py run unchanged.

These stage shims are forward-only (their tensors carry no grad): training goes through the fused step
(NerfTrainer / tgtc_train_step, train.py), stylised rendering through NerfRenderer.render_style.
"""
import torch

from . import _lib


def make_callables(renderer):
    r = renderer

    def sampling_pts_uniform(rays_o, rays_d, N_samples=64, near=0., far=1.05, harmony=False, perturb=False):
        """utils.py:509-531."""
        if harmony:
            raise NotImplementedError("harmony=True (utils.py:516) is never used by the reference's loops")
        rand = None
        if perturb:  # same generator call as utils.py:519-520
            rand = torch.zeros([rays_o.shape[0], N_samples], device=rays_o.device)
            torch.nn.init.uniform_(rand, 0, 1)
        return r.sample_uniform(rays_o, rays_d, N_samples, near, far, rand=rand)

    def sampling_pts_fine_torch(rays_o, rays_d, ts, weights, N_samples_fine=64):
        """utils.py:573-580."""
        return r.sample_fine(rays_o, rays_d, ts, weights, N_samples_fine)

    def alpha_composition(pts_rgb, pts_sigma, t_values, sigma_noise_std=0., white_bkgd=False):
        """utils.py:354-386.  Returns (rgb_exp, t_exp, weights) like the reference (acc is dropped there)."""
        noise = None
        if sigma_noise_std > 0:  # same generator call as utils.py:373-374
            noise = torch.randn(pts_sigma.shape, device=pts_sigma.device) * sigma_noise_std
        rgb, depth, weights, _ = r.composite(pts_rgb, pts_sigma, t_values, noise=noise, white_bkgd=white_bkgd)
        return rgb, depth, weights

    def _forward(which):
        def model_forward(**kwargs):
            """batchify(lambda **kw: model(**kw), chunk)(pts=..., dirs=...)  (utils.py:435-456, models.py:216-223)."""
            r.refresh_weights()
            return r.nerf_forward(which, kwargs["pts"], kwargs["dirs"], want_features=True)
        return model_forward

    return {
        "sampling_pts_uniform": sampling_pts_uniform,
        "sampling_pts_fine_torch": sampling_pts_fine_torch,
        "alpha_composition": alpha_composition,
        "model_forward": _forward(_lib.NET_COARSE),
        "model_forward_fine": _forward(_lib.NET_FINE),
    }


def patch(renderer, modules):
    """Rebinds sampling_pts_uniform / sampling_pts_fine_torch / alpha_composition in each given module
    (the reference's `utils`, `rendering`, `train_tgtcs`).  Returns the callables dict; the two
    model_forward entries are what train() should pass instead of its batchify wrappers."""
    fns = make_callables(renderer)
    for m in modules:
        for name in ("sampling_pts_uniform", "sampling_pts_fine_torch", "alpha_composition"):
            if hasattr(m, name):
                setattr(m, name, fns[name])
    return fns
