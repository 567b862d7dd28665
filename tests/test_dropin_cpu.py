"""CPU: the reference's OWN caller loop (rendering.cal_geometry, imported from the reference tree / its staged copy) driven
through the drop-in seam -- tgtc_style_b200.patch() rebinding sampling_pts_uniform, sampling_pts_fine_torch, alpha_composition
and batchify in the reference's module globals -- against a stand-in renderer that answers the NerfRenderer stage methods with
the oracle.  What this pins without a GPU: the call shapes, keyword names, return structures and the module-recognising
batchify, i.e. that train_tgtcs.train() / rendering.cal_geometry need zero edits.  (The GPU twin runs the same loop on the
kernels: tests/test_gpu_dropin.py.)"""
import numpy as np
import pytest
import torch

import ref_import
import render_oracle as O
import tgtc_style_b200 as T
from tgtc_style_b200 import _lib
from dropin_common import Args, FakeDataset, FakeLoader, POSES

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="needs the reference modules (tree or staged copy)")


class OracleRenderer:
    """NerfRenderer's stage methods answered by the CPU oracle (test infrastructure: checks the seam, not the kernels)"""
    mode = _lib.MLP_FP32

    def __init__(self):
        self._weights_src = [None, None]
        self.calls = []

    def set_weights(self, coarse=None, fine=None):
        for i, src in ((0, coarse), (1, fine)):
            if src is not None:
                self._weights_src[i] = src

    def refresh_weights(self):
        pass

    def _sd(self, net):
        src = self._weights_src[net]
        return {k: v.detach() for k, v in (src.state_dict() if hasattr(src, "state_dict") else src).items()}

    def sample_uniform(self, rays_o, rays_d, n_samples=64, near=0., far=1.05, rand=None, want_pts=True, harmony=False):
        assert not harmony
        self.calls.append("sample_uniform")
        return O.sample_uniform(rays_o, rays_d, n_samples, near, far, rand=rand)

    def sample_fine(self, rays_o, rays_d, ts, weights, n_fine=64):
        self.calls.append("sample_fine")
        return O.sample_fine(rays_o, rays_d, ts, weights, n_fine)

    def nerf_forward(self, net, pts, dirs, want_features=True, mode=None):
        self.calls.append("nerf_forward%d" % net)
        with torch.no_grad():
            return dict(O.nerf_forward(self._sd(net), pts, dirs))

    def composite(self, pts_rgb=None, pts_sigma=None, t_values=None, noise=None, white_bkgd=False, rgbsigma=None):
        self.calls.append("composite")
        rgb, depth, w, acc = O.alpha_composition(pts_rgb, pts_sigma, t_values, noise=noise, white_bkgd=white_bkgd)
        return rgb, depth, w, acc


def test_reference_cal_geometry_runs_unchanged_through_the_seam(tmp_path):
    utils, models, _, _ = ref_import.import_reference()
    rendering = ref_import.import_rendering()
    saved = {(m, n): getattr(m, n) for m in (utils, rendering) for n in T.shims.PATCHED_NAMES}
    r = OracleRenderer()
    try:
        T.patch(r, [utils, rendering])
        for n in T.shims.PATCHED_NAMES:            # star-import copies: both modules' names are rebound
            assert getattr(rendering, n).__self__.__class__ is T.shims.Shims and getattr(utils, n).__self__ is getattr(rendering, n).__self__
        mc, mf = ref_import.reference_nets(models, 0)
        # exactly train_tgtcs.py:30,:37 -- the lambdas close over the modules; `batchify` is the name rendering.py star-imported
        model_forward = rendering.batchify(lambda **kwargs: mc(**kwargs), Args.chunk)
        model_forward_fine = rendering.batchify(lambda **kwargs: mf(**kwargs), Args.chunk)
        H, W, f = 12, 16, 13.0
        ds = FakeDataset(H, W, f, POSES)
        with torch.no_grad():
            rgb_map, t_map = rendering.cal_geometry(model_forward, rendering.sampling_pts_uniform, FakeLoader(ds, 50), Args, torch.device("cpu"),
                                                    sv_path=str(tmp_path), model_forward_fine=model_forward_fine,
                                                    samp_func_fine=rendering.sampling_pts_fine_torch)
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
    assert rgb_map.shape == (2, H, W, 3) and t_map.shape == (2, H, W, 1)
    assert r._weights_src[0] is mc and r._weights_src[1] is mf            # first wrapped = coarse, second = fine
    assert "nerf_forward0" in r.calls and "nerf_forward1" in r.calls and "sample_fine" in r.calls
    ref = O.render_chain(mc.state_dict(), mf.state_dict(), ds.rays_o, ds.rays_d, 0., 1., 64, 64, 1024)
    np.testing.assert_allclose(rgb_map.reshape(-1, 3), ref["rgb"].numpy(), atol=1e-6)
    np.testing.assert_allclose(t_map.reshape(-1), ref["depth"].numpy(), atol=1e-6)
    g = np.load(str(tmp_path / "geometry.npz"))
    assert g["coor_map"].shape == (2, H, W, 3)


def test_batchify_leaves_other_callables_to_the_reference_chunk_loop():
    s = T.shims.Shims(OracleRenderer())
    lin = torch.nn.Linear(3, 2)
    f = s.batchify(lambda **kw: {"y": lin(kw["x"])}, 4)
    x = torch.randn(10, 3)
    assert torch.allclose(f(x=x)["y"], lin(x))
    assert s.batchify(len, None) is len


def test_lazy_return_dict_behaves_like_the_reference_dict():
    calls = []

    def feats():
        calls.append(1)
        return {"base_remap": 1, "pts": 2, "dirs": 3}
    d = T.shims._LazyRet({"rgb": 0, "sigma": 4}, feats)
    assert d["rgb"] == 0 and "base_remap" in d and not calls           # nothing 256-wide is computed until someone reads it
    assert d["base_remap"] == 1 and calls == [1]
    assert set(d) == {"rgb", "sigma", "base_remap", "pts", "dirs"} and calls == [1]
    with pytest.raises(KeyError):
        d["nope"]
