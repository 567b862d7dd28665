"""Host side of the B200 ray-render path: a thin Python layer over the C ABI.

`NerfRenderer` owns one library context (one per device/process) and exposes
  * the fused operator   render(rays_o, rays_d, near, far, chunk) -> {rgb, depth, acc, weights}
    (= the loop body of the reference's rendering.py:27-51),
  * the stage operators that mirror the reference's four injected callables
    (utils.sampling_pts_uniform, models.StyleNerf.forward via utils.batchify,
    utils.alpha_composition, utils.sampling_pts_fine_torch) and its ray generation
    (dataset.get_rays_np + ndc_rays_np).
torch supplies device memory and the stream; all arithmetic happens in
libtgtc_b200.so.  No CPU fallback exists.
"""
import ctypes

import numpy as np
import torch

from . import _lib

LAYER_NAMES = (["net.base_layers.%d" % i for i in range(8)]
               + ["net.sigma_layer", "net.base_remap_layer", "net.rgb_layers.0", "net.rgb_layers.1"])
LAYER_SHAPES = ([(256, 63)] + [(256, 256)] * 4 + [(256, 319)] + [(256, 256)] * 2
                + [(1, 256), (256, 256), (128, 283), (3, 128)])

_MODES = {"fp32": _lib.MLP_FP32, "bf16": _lib.MLP_BF16, "f16": _lib.MLP_F16, "fp16": _lib.MLP_F16,
          _lib.MLP_FP32: _lib.MLP_FP32, _lib.MLP_BF16: _lib.MLP_BF16, _lib.MLP_F16: _lib.MLP_F16}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dptr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_lib.c_double_p)


class NerfRenderer:
    """B200 drop-in for the NeRF ray-render chain of TGTC-Style (coarse + fine StyleNerf)."""

    def __init__(self, device=None, mode="bf16"):
        if not torch.cuda.is_available():
            raise _lib.TgtcError("tgtc-style_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mode = _MODES[mode]
        h = ctypes.c_void_p()
        _lib.check(self.lib.tgtc_create(self.device.index, ctypes.byref(h)))
        self._h = h
        self._ws = None
        self._weights_src = [None, None]   # keeps packed-from tensors alive / version-tracked
        self._versions = [None, None]
        self._keep = [None, None]

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tgtc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    @property
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t, dtype=torch.float32):
        t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != dtype:
            t = t.to(device=self.device, dtype=dtype)
        return t.contiguous()

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def launch_count(self):
        return int(self.lib.tgtc_launch_count(self._h))

    def profile_enable(self, on=True):
        """bracket every MLP launch with CUDA events on its stream (see tgtc_profile_enable)."""
        _lib.check(self.lib.tgtc_profile_enable(self._h, int(on)))

    def profile_read(self):
        """-> (mlp launches, summed device ms, summed algorithmic FLOPs) since the last read."""
        n, ms, fl = _lib.c_i64(0), ctypes.c_double(0), ctypes.c_double(0)
        _lib.check(self.lib.tgtc_profile_read(self._h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)))
        return n.value, ms.value, fl.value

    def profile_read_kind(self, kind):
        """-> (launches, summed device ms, algorithmic FLOPs) of kernel kind 0 fwd / 1 fwd-train / 2 dgrad / 3 wgrad."""
        n, ms, fl = _lib.c_i64(0), ctypes.c_double(0), ctypes.c_double(0)
        _lib.check(self.lib.tgtc_profile_read_kind(self._h, int(kind), ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)))
        return n.value, ms.value, fl.value

    # ------------------------------------------------------------------ weights
    def set_weights(self, coarse=None, fine=None):
        """coarse / fine: a models.StyleNerf (or any nn.Module / state_dict) whose parameters are named
        net.base_layers.{0..7}, net.sigma_layer, net.base_remap_layer, net.rgb_layers.{0,1} (models.py:75-91)."""
        for which, src in ((_lib.NET_COARSE, coarse), (_lib.NET_FINE, fine)):
            if src is None:
                continue
            sd = src.state_dict() if hasattr(src, "state_dict") else src
            tensors = []
            for name, (o, i) in zip(LAYER_NAMES, LAYER_SHAPES):
                w = sd[name + ".weight"]
                b = sd[name + ".bias"]
                if tuple(w.shape) != (o, i) or tuple(b.shape) != (o,):
                    raise ValueError("%s has shape %s/%s, expected %s/%s" % (name, tuple(w.shape), tuple(b.shape), (o, i), (o,)))
                tensors += [self._dev(w.detach()), self._dev(b.detach())]
            arr = (ctypes.c_void_p * _lib.NUM_PARAMS)(*[t.data_ptr() for t in tensors])
            _lib.check(self.lib.tgtc_set_weights(self._h, which, arr, self._stream))
            self._weights_src[which] = src
            self._versions[which] = self._version_of(src)
            self._keep[which] = tensors

    @staticmethod
    def _version_of(src):
        if hasattr(src, "parameters"):
            return tuple(p._version for p in src.parameters())
        return None

    def refresh_weights(self):
        """Re-pack any nn.Module source whose parameters changed in place (optimizer step, load_state_dict)."""
        for which in (0, 1):
            src = self._weights_src[which]
            if src is not None and hasattr(src, "parameters") and self._version_of(src) != self._versions[which]:
                self.set_weights(**{("coarse" if which == 0 else "fine"): src})

    # ------------------------------------------------------------------ K1
    def raygen(self, H, W, K, c2w, ndc=True, ndc_near=1.0, pix_begin=0, n=None, pixel_alignment=False):
        """dataset.get_rays_np + ndc_rays_np (dataset.py:33-61) for pixels [pix_begin, pix_begin+n) -> fp32 [n,3] x2."""
        n = H * W - pix_begin if n is None else n
        ro = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        rd = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        Ka, Kp = _dptr(np.asarray(K, np.float64).reshape(9))
        Ca, Cp = _dptr(np.asarray(c2w, np.float64)[:3, :4].reshape(12))
        _lib.check(self.lib.tgtc_raygen(self._h, H, W, Kp, Cp, int(ndc), float(ndc_near), int(pixel_alignment), pix_begin, n,
                                        _ptr(ro), _ptr(rd), self._stream))
        return ro, rd

    # ------------------------------------------------------------------ K2
    def sample_uniform(self, rays_o, rays_d, n_samples=64, near=0., far=1.05, rand=None, want_pts=True, harmony=False):
        """utils.sampling_pts_uniform (utils.py:509-531).  Returns (pts [N,S,3], ts [N,S])."""
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        n = ro.shape[0]
        ts = torch.empty(n, n_samples, dtype=torch.float32, device=self.device)
        pts = torch.empty(n, n_samples, 3, dtype=torch.float32, device=self.device) if want_pts else None
        rnd = self._dev(rand) if rand is not None else None
        _lib.check(self.lib.tgtc_sample_uniform(self._h, _ptr(ro), _ptr(rd), n, n_samples, float(near), float(far), int(bool(harmony)),
                                                _ptr(rnd), _ptr(pts), _ptr(ts), self._stream))
        return pts, ts

    # ------------------------------------------------------------------ K3+K4
    def nerf_forward(self, net, pts, dirs, want_features=True, mode=None):
        """models.StyleNerf.forward(pts=[N,S,3], dirs=[N,S,3]) through utils.batchify (models.py:216-223).
        Returns the reference's dict: rgb, base_remap, pts (embedded), sigma, dirs (embedded)."""
        mode = self.mode if mode is None else _MODES[mode]
        pts_t = torch.as_tensor(pts)
        dirs_t = torch.as_tensor(dirs)
        if pts_t.dim() != 3 or pts_t.shape[-1] != 3:
            raise ValueError("pts must be [N,S,3]")
        n, S = pts_t.shape[0], pts_t.shape[1]
        # the reference passes rays_d.unsqueeze(1).expand(N,S,3) (rendering.py:30): a stride-0 view
        per_ray = dirs_t.dim() == 2 or (dirs_t.dim() == 3 and (dirs_t.stride(1) == 0 or S == 1))
        d = self._dev(dirs_t[:, 0, :] if (per_ray and dirs_t.dim() == 3) else dirs_t)
        p = self._dev(pts_t)
        rgb = torch.empty(n, S, 3, dtype=torch.float32, device=self.device)
        sigma = torch.empty(n, S, dtype=torch.float32, device=self.device)
        if want_features and mode != _lib.MLP_FP32:
            mode = _lib.MLP_FP32  # feature outputs exist on the general path only
        remap = torch.empty(n, S, 256, dtype=torch.float32, device=self.device) if want_features else None
        pe = torch.empty(n, S, 63, dtype=torch.float32, device=self.device) if want_features else None
        de = torch.empty(n, S, 27, dtype=torch.float32, device=self.device) if want_features else None
        _lib.check(self.lib.tgtc_nerf_forward(self._h, net, mode, _ptr(p), _ptr(d), int(per_ray), n, S, _ptr(rgb), _ptr(sigma),
                                              _ptr(remap), _ptr(pe), _ptr(de), self._stream))
        out = {"rgb": rgb, "sigma": sigma}
        if want_features:
            out.update(base_remap=remap, pts=pe, dirs=de)
        return out

    def nerf_forward_rays(self, net, rays_o, rays_d, ts=None, n_samples=64, near=0., far=1., mode=None):
        """Fused K2+K3+K4: rgbsigma [N,S,4] for samples o + t*d (ts=None -> uniform coarse positions)."""
        mode = self.mode if mode is None else _MODES[mode]
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        n = ro.shape[0]
        tsd = self._dev(ts) if ts is not None else None
        S = tsd.shape[1] if tsd is not None else n_samples
        out = torch.empty(n, S, 4, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_nerf_forward_rays(self._h, net, mode, _ptr(ro), _ptr(rd), _ptr(tsd), n, S, float(near), float(far),
                                                   _ptr(out), self._stream))
        return out

    # stage-level training entries (the autograd.Function behind the model_forward drop-in, shims.py)
    def nerf_forward_stash(self, net, pts, dirs):
        """Training forward of one net on explicit points: pts [N,S,3], dirs [N,3] -> (rgbsigma [N,S,4], stash).  The stash (a uint8
        device tensor owned by the caller) holds the activations nerf_backward needs; S in {64,128}; bf16 tcgen05 kernels."""
        p, d = self._dev(pts), self._dev(dirs)
        n, S = p.shape[0], p.shape[1]
        sb = int(self.lib.tgtc_nerf_stash_bytes(n, S))
        stash = torch.empty(sb + 1024, dtype=torch.uint8, device=self.device)
        off = (-stash.data_ptr()) % 1024
        rs = torch.empty(n, S, 4, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_nerf_forward_stash(self._h, net, _ptr(p), _ptr(d), n, S, _ptr(rs), ctypes.c_void_p(stash.data_ptr() + off), sb,
                                                    self._stream))
        return rs, (stash, off, sb)

    def nerf_backward(self, net, dirs, rgbsigma, d_rgbsigma, stash, grads=None, accumulate=False):
        """dL/d(r,g,b,sigma) [N,S,4] -> flat fp32 gradient [tgtc_num_params()] of net `net` (tgtc_set_weights order)."""
        d, rs, drs = self._dev(dirs), self._dev(rgbsigma), self._dev(d_rgbsigma)
        n, S = rs.shape[0], rs.shape[1]
        P = int(self.lib.tgtc_num_params())
        if grads is None:
            grads = torch.empty(P, dtype=torch.float32, device=self.device)
            accumulate = False
        scb = int(self.lib.tgtc_nerf_backward_scratch_bytes(self._h, n, S))
        ws = self._workspace(scb + 1024)
        woff = (-ws.data_ptr()) % 1024
        buf, off, sb = stash
        _lib.check(self.lib.tgtc_nerf_backward(self._h, net, _ptr(d), n, S, _ptr(rs), _ptr(drs), _ptr(grads), int(accumulate),
                                               ctypes.c_void_p(buf.data_ptr() + off), sb, ctypes.c_void_p(ws.data_ptr() + woff), scb,
                                               self._stream))
        return grads

    # ------------------------------------------------------------------ K5
    def composite(self, pts_rgb=None, pts_sigma=None, t_values=None, noise=None, white_bkgd=False, rgbsigma=None):
        """utils.alpha_composition (utils.py:354-386).  Returns (rgb [N,3], depth [N], weights [N,S], acc [N])."""
        ts_t = torch.as_tensor(t_values)
        if rgbsigma is not None:
            rs = self._dev(rgbsigma)
            n, S = rs.shape[0], rs.shape[1]
            rgb = sig = None
        else:
            rgb, sig = self._dev(pts_rgb), self._dev(pts_sigma)
            n, S = sig.shape[0], sig.shape[1]
            rs = None
        if ts_t.dim() == 2 and ts_t.shape[0] == n and ts_t.stride(0) == 0 and n > 1:
            tsd, stride = self._dev(ts_t[0]), 0     # the reference's expanded coarse ts (utils.py:512)
        elif ts_t.dim() == 1:
            tsd, stride = self._dev(ts_t), 0
        else:
            tsd, stride = self._dev(ts_t), S
        nz = self._dev(noise) if noise is not None else None
        o_rgb = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        o_d = torch.empty(n, dtype=torch.float32, device=self.device)
        o_a = torch.empty(n, dtype=torch.float32, device=self.device)
        o_w = torch.empty(n, S, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_composite(self._h, _ptr(rgb), _ptr(sig), _ptr(rs), _ptr(tsd), stride, _ptr(nz), int(white_bkgd), n, S,
                                           _ptr(o_rgb), _ptr(o_d), _ptr(o_a), _ptr(o_w), self._stream))
        return o_rgb, o_d, o_w, o_a

    def composite_backward(self, rgbsigma, t_values, g_rgb, g_depth=None, g_acc=None, noise=None, white_bkgd=False):
        """Gradient of utils.alpha_composition (utils.py:354-386) w.r.t. the per-sample (r,g,b,sigma): rgbsigma [N,S,4],
        g_rgb [N,3] = dL/d(rgb map) (+ optional dL/d(depth), dL/d(acc)) -> d_rgbsigma [N,S,4]."""
        rs = self._dev(rgbsigma)
        n, S = rs.shape[0], rs.shape[1]
        ts_t = torch.as_tensor(t_values)
        if ts_t.dim() == 1 or (ts_t.dim() == 2 and ts_t.stride(0) == 0 and n > 1):
            tsd, stride = self._dev(ts_t if ts_t.dim() == 1 else ts_t[0]), 0
        else:
            tsd, stride = self._dev(ts_t), S
        g = self._dev(g_rgb)
        gd = self._dev(g_depth) if g_depth is not None else None
        ga = self._dev(g_acc) if g_acc is not None else None
        nz = self._dev(noise) if noise is not None else None
        out = torch.empty(n, S, 4, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_composite_backward(self._h, _ptr(rs), _ptr(tsd), stride, _ptr(nz), int(white_bkgd), n, S, _ptr(g),
                                                    _ptr(gd), _ptr(ga), _ptr(out), self._stream))
        return out

    # ------------------------------------------------------------------ K6+K7
    def sample_fine(self, rays_o, rays_d, ts, weights, n_fine=64, want_pts=True, return_aux=False):
        """utils.sampling_pts_fine_torch (utils.py:573-580).  Returns (pts [N,S+F,3], ts [N,S+F]) and, with
        return_aux, also (t_samples [N,F], inds [N,F] int64)."""
        w = self._dev(weights)
        n, S = w.shape
        ts_t = torch.as_tensor(ts)
        if ts_t.dim() == 1 or (ts_t.stride(0) == 0 and n > 1):
            tsd, stride = self._dev(ts_t if ts_t.dim() == 1 else ts_t[0]), 0
        else:
            tsd, stride = self._dev(ts_t), S
        T = S + n_fine
        ts_out = torch.empty(n, T, dtype=torch.float32, device=self.device)
        ro = self._dev(rays_o) if want_pts else None
        rd = self._dev(rays_d) if want_pts else None
        pts = torch.empty(n, T, 3, dtype=torch.float32, device=self.device) if want_pts else None
        inds = torch.empty(n, n_fine, dtype=torch.int64, device=self.device) if return_aux else None
        smp = torch.empty(n, n_fine, dtype=torch.float32, device=self.device) if return_aux else None
        _lib.check(self.lib.tgtc_sample_fine(self._h, _ptr(ro), _ptr(rd), _ptr(tsd), stride, _ptr(w), n, S, n_fine, _ptr(pts),
                                             _ptr(ts_out), _ptr(inds), _ptr(smp), self._stream))
        if return_aux:
            return pts, ts_out, smp, inds
        return pts, ts_out

    # ------------------------------------------------------------------ the fused operator
    def _alloc_out(self, n, S, F, extras, device, pin=False):
        kw = dict(dtype=torch.float32, device=device)
        if pin:
            kw["pin_memory"] = True
        T = S + F
        out = {"rgb": torch.empty(n, 3, **kw), "depth": torch.empty(n, **kw), "acc": torch.empty(n, **kw),
               "weights": torch.empty(n, T, **kw)}
        if extras:
            out.update(rgb_coarse=torch.empty(n, 3, **kw), depth_coarse=torch.empty(n, **kw), acc_coarse=torch.empty(n, **kw),
                       weights_coarse=torch.empty(n, S, **kw), ts_fine=torch.empty(n, T, **kw))
        return out

    @staticmethod
    def _out_struct(out):
        s = _lib.RenderOut()
        for name in _lib.OUT_FIELDS:
            t = out.get(name)
            setattr(s, name, t.data_ptr() if t is not None else None)
        return s

    def render(self, rays_o, rays_d, near=0., far=1., chunk=None, n_samples=64, n_fine=64, white_bkgd=False, extras=False,
               want_weights=True, mode=None, out=None):
        """render(rays_o [N,3], rays_d [N,3], near, far, chunk) -> {rgb [N,3], depth [N], acc [N], weights [N,S+F]}
        (+ rgb_coarse, depth_coarse, acc_coarse, weights_coarse, ts_fine when extras).  Device tensors in and out."""
        self.refresh_weights()
        mode = self.mode if mode is None else _MODES[mode]
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        n = ro.shape[0]
        if out is None:
            out = self._alloc_out(n, n_samples, n_fine, extras, self.device)
            if not want_weights:
                out.pop("weights")
        ck = int(chunk) if chunk else 0
        wsb = self.lib.tgtc_render_workspace_bytes_mode(mode, n, n_samples, n_fine, ck)
        ws = self._workspace(wsb)
        s = self._out_struct(out)
        _lib.check(self.lib.tgtc_render(self._h, mode, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine, ck,
                                        int(white_bkgd), ctypes.byref(s), _ptr(ws), wsb, self._stream))
        return out

    def render_host(self, rays_o, rays_d, near=0., far=1., chunk=None, n_samples=64, n_fine=64, white_bkgd=False, extras=False,
                    want_weights=False, mode=None, out=None):
        """Same operator with HOST tensors (pinned recommended): H2D of the rays, render, D2H of the results,
        synchronous.  This is the call the reference's loop makes per batch (rendering.py:20-21, :53-54)."""
        self.refresh_weights()
        mode = self.mode if mode is None else _MODES[mode]
        ro = torch.as_tensor(rays_o, dtype=torch.float32).contiguous()
        rd = torch.as_tensor(rays_d, dtype=torch.float32).contiguous()
        if ro.is_cuda or rd.is_cuda:
            raise ValueError("render_host takes host tensors")
        n = ro.shape[0]
        if out is None:
            out = self._alloc_out(n, n_samples, n_fine, extras, "cpu", pin=True)
            if not want_weights:
                out.pop("weights")
        s = self._out_struct(out)
        _lib.check(self.lib.tgtc_render_host(self._h, mode, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine,
                                             int(chunk) if chunk else 0, int(white_bkgd), ctypes.byref(s), self._stream))
        return out

    def render_frame(self, H, W, K, c2w, pix_begin=0, n=None, near=0., far=1., chunk=None, n_samples=64, n_fine=64, ndc=True,
                     ndc_near=1.0, white_bkgd=False, extras=False, want_weights=False, mode=None, out=None, pixel_alignment=False):
        """Ray generation (K1) fused in front of render() for pixels [pix_begin, pix_begin+n) of an HxW frame."""
        self.refresh_weights()
        mode = self.mode if mode is None else _MODES[mode]
        n = H * W - pix_begin if n is None else n
        if out is None:
            out = self._alloc_out(n, n_samples, n_fine, extras, self.device)
            if not want_weights:
                out.pop("weights")
        ck = int(chunk) if chunk else 0
        wsb = self.lib.tgtc_render_frame_workspace_bytes_mode(mode, n, n_samples, n_fine, ck)
        ws = self._workspace(wsb)
        Ka, Kp = _dptr(np.asarray(K, np.float64).reshape(9))
        Ca, Cp = _dptr(np.asarray(c2w, np.float64)[:3, :4].reshape(12))
        s = self._out_struct(out)
        _lib.check(self.lib.tgtc_render_frame(self._h, mode, H, W, Kp, Cp, int(ndc), float(ndc_near), int(bool(pixel_alignment)), pix_begin, n, float(near),
                                              float(far), n_samples, n_fine, ck, int(white_bkgd), ctypes.byref(s), _ptr(ws), wsb,
                                              self._stream))
        return out

    # ------------------------------------------------------------------ per-ray style head (f1)
    STYLE_C_SHAPES = [(256, 95), (256, 288), (256, 288), (256, 288), (256, 351)]
    STYLE_W_SHAPES = [(256, 607), (256, 288), (256, 288), (256, 288), (256, 351), (256, 288), (256, 288), (3, 288)]

    def set_style_weights(self, concat_style, style):
        """concat_style: models.StyleMLP_before_concat, style: models.StyleMLP_Wild_multilayers (nn.Modules or state_dicts with
        keys layers.{i}.weight / layers.{i}.bias; models.py:120-180)."""
        cached = getattr(self, "_style_src", None)
        if cached is not None and cached[0] is concat_style and cached[1] is style and isinstance(style, dict):
            # the same two dicts of device tensors as last time (a trainer re-packing after an optimizer step): skip the checks
            _lib.check(self.lib.tgtc_set_style_weights(self._h, self._style_arr, self._stream))
            return
        tensors = []
        for src, shapes in ((concat_style, self.STYLE_C_SHAPES), (style, self.STYLE_W_SHAPES)):
            sd = src.state_dict() if hasattr(src, "state_dict") else src
            for i, (o, k) in enumerate(shapes):
                w, b = sd["layers.%d.weight" % i], sd["layers.%d.bias" % i]
                if tuple(w.shape) != (o, k) or tuple(b.shape) != (o,):
                    raise ValueError("style layers.%d has shape %s/%s, expected %s/%s" % (i, tuple(w.shape), tuple(b.shape), (o, k), (o,)))
                tensors += [self._dev(w.detach()), self._dev(b.detach())]
        arr = (ctypes.c_void_p * 26)(*[t.data_ptr() for t in tensors])
        _lib.check(self.lib.tgtc_set_style_weights(self._h, arr, self._stream))
        self._style_keep = tensors
        on_device = all(t.data_ptr() == sd_t.data_ptr() for t, sd_t in zip(tensors[:2], (concat_style["layers.0.weight"], concat_style["layers.0.bias"]))) \
            if isinstance(concat_style, dict) else False
        self._style_src, self._style_arr = ((concat_style, style), arr) if on_device else (None, None)

    def render_style(self, rays_o, rays_d, latents, near=0., far=1., chunk=None, n_samples=64, n_fine=64, extras=False,
                     want_weights=False, out=None, mode=None):
        """The loop body of render_style (rendering.py:118-178, perturb=False) for one batch of rays.
        latents = the output of latents_model_1 (rendering.py:125): [32] for a batch that shares one (style, frame), or [N,32]
        per-ray latents (one library call either way: uniform latents fold into per-call biases, mixed ones into per-ray biases).
        -> {rgb, depth, acc} (+ weights / coarse outputs / ts_fine like render())."""
        self.refresh_weights()
        mode = self.mode if mode is None else _MODES[mode]
        if mode == _lib.MLP_FP32:
            raise _lib.TgtcError("the stylised render runs on the tensor-core path: mode must be 'bf16' or 'f16'")
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        n = ro.shape[0]
        lat = self._dev(latents)
        if out is None:
            out = self._alloc_out(n, n_samples, n_fine, extras, self.device)
            if not want_weights:
                out.pop("weights")
        ck = int(chunk) if chunk else 0
        if lat.dim() == 2 and n > 1 and not bool((lat == lat[:1]).all()):
            # per-ray latents inside ONE library call (tgtc_render_style_rays): the latent columns of every layer become per-ray
            # effective biases; batches that mix (style, frame) pairs need no host-side splitting
            if lat.shape[0] != n:
                raise ValueError("latents must be [32] or [N,32]")
            lat = lat.contiguous()
            wsb = self.lib.tgtc_render_style_rays_workspace_bytes(n, n_samples, n_fine, ck)
            ws = self._workspace(wsb + 1024)
            off = (-ws.data_ptr()) % 1024
            s = self._out_struct(out)
            _lib.check(self.lib.tgtc_render_style_rays(self._h, mode, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine, ck,
                                                       _ptr(lat), ctypes.byref(s), ctypes.c_void_p(ws.data_ptr() + off), wsb, self._stream))
            return out
        if lat.dim() == 2:
            lat = lat[0]
        lat1 = lat.reshape(32).contiguous()
        lat2 = lat1.mean().expand(32).contiguous()          # rendering.py:126: mean over the latent dim, broadcast (:139)
        wsb = self.lib.tgtc_render_style_workspace_bytes(n, n_samples, n_fine, ck)
        ws = self._workspace(wsb + 1024)
        off = (-ws.data_ptr()) % 1024
        s = self._out_struct(out)
        _lib.check(self.lib.tgtc_render_style(self._h, mode, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine, ck,
                                              _ptr(lat1), _ptr(lat2), ctypes.byref(s), ctypes.c_void_p(ws.data_ptr() + off), wsb,
                                              self._stream))
        return out

    def set_coarse_event(self, event):
        """torch.cuda.Event (or None) that every following train_step records once the coarse net's gradient half is final
        (tgtc_train_set_coarse_event): lets the caller all-reduce that half underneath the fine net's work."""
        if event is not None:
            event.record(torch.cuda.current_stream(self.device))      # materialises the lazily-created CUDA event
            handle = ctypes.c_void_p(event.cuda_event)
        else:
            handle = None
        _lib.check(self.lib.tgtc_train_set_coarse_event(self._h, handle))
        self._coarse_event = event

    # stage entries of the style head on explicit features (the injected concat_style_forward / style_forward callables)
    def _style_stage_ws(self, M):
        wsb = int(self.lib.tgtc_style_stage_workspace_bytes(int(M)))
        ws = self._workspace(wsb + 1024)
        return ctypes.c_void_p(ws.data_ptr() + ((-ws.data_ptr()) % 1024)), wsb

    def style_concat_forward(self, x, latent, mode=None):
        """StyleMLP_before_concat.forward(x=[...,63] embedded pts, latent=[32]) -> concat_features [...,256] (models.py:137-147)."""
        mode = self.mode if mode is None else _MODES[mode]
        mode = _lib.MLP_F16 if mode == _lib.MLP_FP32 else mode
        xs = self._dev(x)
        M = xs.numel() // 63
        out = torch.empty(tuple(xs.shape[:-1]) + (256,), dtype=torch.float32, device=self.device)
        lat = self._dev(latent).reshape(32).contiguous()
        ws, wsb = self._style_stage_ws(M)
        _lib.check(self.lib.tgtc_style_concat_forward(self._h, mode, _ptr(xs), _ptr(lat), M, _ptr(out), ws, wsb, self._stream))
        return out

    def style_forward(self, x, concated, latent, mode=None):
        """StyleMLP_Wild_multilayers.forward(x=[...,63], concated=[...,512], latent=[32]) -> rgb [...,3] (models.py:165-180)."""
        mode = self.mode if mode is None else _MODES[mode]
        mode = _lib.MLP_F16 if mode == _lib.MLP_FP32 else mode
        xs, cc = self._dev(x), self._dev(concated)
        M = xs.numel() // 63
        out = torch.empty(tuple(xs.shape[:-1]) + (3,), dtype=torch.float32, device=self.device)
        lat = self._dev(latent).reshape(32).contiguous()
        ws, wsb = self._style_stage_ws(M)
        _lib.check(self.lib.tgtc_style_forward(self._h, mode, _ptr(xs), _ptr(cc), _ptr(lat), M, _ptr(out), ws, wsb, self._stream))
        return out

    # ------------------------------------------------------------------ training step (a11)
    def train_step(self, rays_o, rays_d, rgb_gt, n_total=None, near=0., far=1., n_samples=64, n_fine=64, grads=None,
                   accumulate=False, rand=None, noise_coarse=None, noise_fine=None, seed=None, perturb=False, sigma_noise_std=0.):
        """Forward + backward of Origin_train's loss (train_tgtcs.py:228-255, perturb=0, noise=0) for one batch of rays:
        loss = mse(rgb_gt, rgb_coarse) + mse(rgb_gt, rgb_fine), means taken over n_total rays (default: this batch).
        Returns {"loss" (device scalar), "grads" (flat fp32 [2*P]: coarse net then fine net, tgtc_set_weights order),
        "rgb_coarse", "rgb_fine"}.  Pass the same `grads` with accumulate=True for the later ray chunks of one step.
        rand [n,S] (uniforms, perturb=True of utils.py:518-524) and noise_coarse [n,S] / noise_fine [n,S+F] (randn*std,
        utils.py:372-374) replay the reference's stochastic options with caller-drawn tensors; with `seed` (an int) they are
        drawn inside the kernels instead (Philox4x32-10, tgtc_train_step_seeded): perturb / sigma_noise_std select which."""
        self.refresh_weights()
        ro, rd, gt = self._dev(rays_o), self._dev(rays_d), self._dev(rgb_gt)
        n = ro.shape[0]
        n_total = n if n_total is None else int(n_total)
        P = int(self.lib.tgtc_num_params())
        if grads is None:
            grads = torch.empty(2 * P, dtype=torch.float32, device=self.device)
            accumulate = False
        sums = torch.zeros(2, dtype=torch.float32, device=self.device)
        rgb_c = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        rgb_f = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        wsb = self.lib.tgtc_train_workspace_bytes(self._h, n, n_samples, n_fine)
        ws = self._workspace(wsb + 1024)
        off = (-ws.data_ptr()) % 1024
        rnd = self._dev(rand) if rand is not None else None
        nzc = self._dev(noise_coarse) if noise_coarse is not None else None
        nzf = self._dev(noise_fine) if noise_fine is not None else None
        if seed is not None:
            if rnd is not None or nzc is not None or nzf is not None:
                raise ValueError("pass either replay tensors or a seed")
            _lib.check(self.lib.tgtc_train_step_seeded(self._h, _ptr(ro), _ptr(rd), _ptr(gt), n, n_total, float(near), float(far),
                                                       n_samples, n_fine, int(seed) & 0xFFFFFFFFFFFFFFFF, int(bool(perturb)),
                                                       float(sigma_noise_std), _ptr(grads), int(accumulate), _ptr(sums), _ptr(rgb_c),
                                                       _ptr(rgb_f), ctypes.c_void_p(ws.data_ptr() + off), wsb, self._stream))
        else:
            _lib.check(self.lib.tgtc_train_step(self._h, _ptr(ro), _ptr(rd), _ptr(gt), n, n_total, float(near), float(far), n_samples,
                                                n_fine, _ptr(rnd), _ptr(nzc), _ptr(nzf), _ptr(grads), int(accumulate), _ptr(sums),
                                                _ptr(rgb_c), _ptr(rgb_f),
                                                ctypes.c_void_p(ws.data_ptr() + off), wsb, self._stream))
        return {"loss": sums.sum() / (3.0 * n_total), "grads": grads, "rgb_coarse": rgb_c, "rgb_fine": rgb_f}

    # ------------------------------------------------------------------ Style_train (train_tgtcs.py:311-495)
    def style_train_forward(self, rays_o, rays_d, latents, near=0., far=1., n_samples=64, n_fine=64, rand=None, noise_coarse=None,
                            noise_fine=None, workspace=None, seed=None, perturb=True, sigma_noise_std=0.):
        """Forward of one Style_train batch (train_tgtcs.py:404-479): frozen NeRF nets, both style modules with per-ray
        latents [N,32], stratified positions replaying `rand` [N,S] (perturb=True) -> {"rgb_coarse", "rgb_fine"} plus an
        opaque "state" for style_train_backward (the activation stash lives in `workspace`, a uint8 device tensor; by default
        one owned by this call's state -- use separate workspaces for batches whose backward passes are both pending).
        With `seed` (an int) the jitter (perturb) and sigma noise are drawn inside the kernels (Philox) instead of replayed."""
        self.refresh_weights()
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        lat = self._dev(latents).contiguous()
        n = ro.shape[0]
        if tuple(lat.shape) != (n, 32):
            raise ValueError("latents must be [N,32]")
        wsb = self.lib.tgtc_style_train_workspace_bytes(self._h, n, n_samples, n_fine)
        ws = workspace if workspace is not None else torch.empty(int(wsb) + 1024, dtype=torch.uint8, device=self.device)
        if ws.numel() < wsb + 1024:
            raise ValueError("workspace needs %d bytes" % (wsb + 1024))
        off = (-ws.data_ptr()) % 1024
        rnd = self._dev(rand) if rand is not None else None
        nzc = self._dev(noise_coarse) if noise_coarse is not None else None
        nzf = self._dev(noise_fine) if noise_fine is not None else None
        rgb_c = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        rgb_f = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        seeded = None
        if seed is not None:
            if rnd is not None or nzc is not None or nzf is not None:
                raise ValueError("pass either replay tensors or a seed")
            seeded = (int(seed) & 0xFFFFFFFFFFFFFFFF, int(bool(perturb)), float(sigma_noise_std))
            _lib.check(self.lib.tgtc_style_train_forward_seeded(self._h, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine,
                                                                _ptr(lat), seeded[0], seeded[1], seeded[2], _ptr(rgb_c), _ptr(rgb_f),
                                                                ctypes.c_void_p(ws.data_ptr() + off), wsb, self._stream))
        else:
            _lib.check(self.lib.tgtc_style_train_forward(self._h, _ptr(ro), _ptr(rd), n, float(near), float(far), n_samples, n_fine,
                                                         _ptr(lat), _ptr(rnd), _ptr(nzc), _ptr(nzf), _ptr(rgb_c), _ptr(rgb_f),
                                                         ctypes.c_void_p(ws.data_ptr() + off), wsb, self._stream))
        state = dict(n=n, S=n_samples, F=n_fine, lat=lat, rand=rnd, nzc=nzc, nzf=nzf, ws=ws, off=off, wsb=wsb, seeded=seeded)
        return {"rgb_coarse": rgb_c, "rgb_fine": rgb_f, "state": state}

    def style_train_workspace_bytes(self, n, n_samples=64, n_fine=64):
        return int(self.lib.tgtc_style_train_workspace_bytes(self._h, int(n), n_samples, n_fine)) + 1024

    def style_num_params(self):
        return int(self.lib.tgtc_style_num_params())

    def style_train_backward(self, state, d_rgb_coarse, d_rgb_fine, grads=None, accumulate=False):
        """Backward of the batch `state` came from: dL/d rgb_coarse, dL/d rgb_fine [N,3] -> {"grads": flat fp32
        [tgtc_style_num_params()] in set_style_weights order, "d_latents": [N,32]}."""
        gc, gf = self._dev(d_rgb_coarse).contiguous(), self._dev(d_rgb_fine).contiguous()
        P = int(self.lib.tgtc_style_num_params())
        if grads is None:
            grads = torch.empty(P, dtype=torch.float32, device=self.device)
            accumulate = False
        dlat = torch.empty(state["n"], 32, dtype=torch.float32, device=self.device)
        if state.get("seeded") is not None:
            sd = state["seeded"]
            _lib.check(self.lib.tgtc_style_train_backward_seeded(self._h, state["n"], state["S"], state["F"], _ptr(state["lat"]), sd[0],
                                                                 sd[1], sd[2], _ptr(gc), _ptr(gf), _ptr(grads), int(accumulate), _ptr(dlat),
                                                                 ctypes.c_void_p(state["ws"].data_ptr() + state["off"]), state["wsb"],
                                                                 self._stream))
            return {"grads": grads, "d_latents": dlat}
        _lib.check(self.lib.tgtc_style_train_backward(self._h, state["n"], state["S"], state["F"], _ptr(state["lat"]),
                                                      int(state["rand"] is not None), _ptr(state["nzc"]), _ptr(state["nzf"]), _ptr(gc),
                                                      _ptr(gf), _ptr(grads), int(accumulate), _ptr(dlat),
                                                      ctypes.c_void_p(state["ws"].data_ptr() + state["off"]), state["wsb"], self._stream))
        return {"grads": grads, "d_latents": dlat}

    def style_loss_sums(self, rgb_c, rgb_f, gt, coh=None):
        """-> device tensor [4]: squared-error sums of the two maps and, with coh = (c2, f2, x, y, rgb_origin, prev_origin), the two
        sums of squared cosine-similarity differences of the coherence term (tgtc_style_loss_sums)."""
        sums = torch.empty(4, dtype=torch.float32, device=self.device)
        c = [self._dev(t).contiguous() for t in coh] if coh is not None else [None] * 6
        n2 = c[0].shape[0] if coh is not None else 0
        _lib.check(self.lib.tgtc_style_loss_sums(self._h, _ptr(rgb_c), _ptr(rgb_f), _ptr(gt), rgb_c.shape[0], _ptr(c[0]), _ptr(c[1]),
                                                 _ptr(c[2]), _ptr(c[3]), _ptr(c[4]), _ptr(c[5]), n2, _ptr(sums), self._stream))
        return sums

    def style_loss_grads(self, rgb_c, rgb_f, gt, scale_rgb, coh=None, coh_ss=None, scale_coh=0.0, out=None):
        """-> (d rgb_c, d rgb_f, d c2 or None, d f2 or None): gradients of scale_rgb * (sum sq err) + scale_coh * (the two L2 norms
        built from coh_ss [2], device) (tgtc_style_loss_grads).  out: the four result tensors to write into (contiguous)."""
        n = rgb_c.shape[0]
        d_c, d_f = (out[0], out[1]) if out is not None else (torch.empty_like(rgb_c), torch.empty_like(rgb_f))
        if coh is None:
            c, d2c, d2f, n2 = [None] * 6, None, None, 0
        else:
            c = [self._dev(t).contiguous() for t in coh]
            n2 = c[0].shape[0]
            d2c, d2f = (out[2], out[3]) if out is not None else (torch.empty_like(c[0]), torch.empty_like(c[1]))
        _lib.check(self.lib.tgtc_style_loss_grads(self._h, _ptr(rgb_c), _ptr(rgb_f), _ptr(gt), n, _ptr(c[0]), _ptr(c[1]), _ptr(c[2]),
                                                  _ptr(c[3]), _ptr(c[4]), _ptr(c[5]), n2, _ptr(coh_ss), float(scale_rgb),
                                                  float(scale_coh), _ptr(d_c), _ptr(d_f), _ptr(d2c), _ptr(d2f), self._stream))
        return d_c, d_f, d2c, d2f

    def style_latents_forward(self, table, mu, logvar, style_id, frame_id, n_logp, frame_num, sigma_scale=1.0, table_tiles=7):
        """models.StyleLatents_variational.forward for every ray + the minus_logp sum of the first n_logp rays
        (tgtc_style_latents_forward) -> (lat [n,32], logp_sum [1])."""
        n = style_id.shape[0]
        lat = torch.empty(n, 32, dtype=torch.float32, device=self.device)
        logp = torch.empty(1, dtype=torch.float32, device=self.device)
        rows = table.numel() // 32
        _lib.check(self.lib.tgtc_style_latents_forward(self._h, _ptr(table), _ptr(mu), _ptr(logvar), _ptr(style_id), _ptr(frame_id), n,
                                                       int(n_logp), rows, int(frame_num), int(table_tiles), float(sigma_scale), _ptr(lat), _ptr(logp),
                                                       self._stream))
        return lat, logp

    def style_latents_backward(self, table, mu, logvar, style_id, frame_id, n_logp, frame_num, dlat, logp_scale, sigma_scale=1.0,
                               out=None, accumulate=False, table_tiles=7):
        """gradient of <dlat, lat> + logp_scale * logp_sum w.r.t. the table (tgtc_style_latents_backward) -> tensor like table."""
        n = style_id.shape[0]
        if out is None:
            out = torch.empty_like(table)
            accumulate = False
        rows = table.numel() // 32
        _lib.check(self.lib.tgtc_style_latents_backward(self._h, _ptr(table), _ptr(mu), _ptr(logvar), _ptr(style_id), _ptr(frame_id), n,
                                                        int(n_logp), rows, int(frame_num), int(table_tiles), float(sigma_scale), _ptr(dlat),
                                                        float(logp_scale), _ptr(out), int(accumulate), self._stream))
        return out

    def style_grad_views(self, flat):
        """Per-parameter views into a flat style gradient buffer: (concat-module dict, wild-module dict), state_dict keys."""
        out, o = [], 0
        for shapes in (self.STYLE_C_SHAPES, self.STYLE_W_SHAPES):
            d = {}
            for i, (no, ni) in enumerate(shapes):
                d["layers.%d.weight" % i] = flat[o:o + no * ni].view(no, ni)
                o += no * ni
                d["layers.%d.bias" % i] = flat[o:o + no]
                o += no
            out.append(d)
        return tuple(out)

    def philox_fill(self, seed, stream_id, n, normal=False, std=1.0):
        """The tensor a seeded training step draws in-kernel: stream 0 = jitter uniforms, 1 / 2 = coarse / fine sigma noise."""
        out = torch.empty(int(n), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_philox_fill(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id), int(bool(normal)), float(std), int(n),
                                             _ptr(out), self._stream))
        return out

    def adam_step(self, params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        """torch.optim.Adam's update (train_tgtcs.py:39) on flat fp32 device buffers, one kernel (tgtc_adam_step)."""
        n = params.numel()
        _lib.check(self.lib.tgtc_adam_step(self._h, _ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), n, float(lr),
                                           float(betas[0]), float(betas[1]), float(eps), int(step), self._stream))

    def grad_views(self, flat):
        """Per-parameter views into a flat gradient buffer: (coarse dict, fine dict) keyed like the state_dict."""
        P = int(self.lib.tgtc_num_params())
        out = []
        for net in range(2):
            d, o = {}, net * P
            for name, (no, ni) in zip(LAYER_NAMES, LAYER_SHAPES):
                d[name + ".weight"] = flat[o:o + no * ni].view(no, ni)
                o += no * ni
                d[name + ".bias"] = flat[o:o + no]
                o += no
            out.append(d)
        return tuple(out)

    # ------------------------------------------------------------------ test hook
    def debug_tc_layers(self, net, rays_o, rays_d, ts, n_samples, near, far, layers):
        self.lib.tgtc_debug_tc_f16(int(self.mode == _lib.MLP_F16))
        ro, rd = self._dev(rays_o), self._dev(rays_d)
        n = ro.shape[0]
        tsd = self._dev(ts) if ts is not None else None
        S = tsd.shape[1] if tsd is not None else n_samples
        acc = torch.zeros(n * S, 256, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.tgtc_debug_tc_layers(self._h, net, _ptr(ro), _ptr(rd), _ptr(tsd), n, S, float(near), float(far), layers,
                                                 _ptr(acc), self._stream))
        return acc
