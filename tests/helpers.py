"""Shared test helpers: golden fixtures, synthetic inputs, and a bf16-operand emulation of the tcgen05 MLP."""
import functools
import os

import numpy as np
import torch

import render_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FERN_SMALL = (378, 504, 407.566)
FERN_FULL = (756, 1008, 815.13)


@functools.lru_cache(maxsize=None)
def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@functools.lru_cache(maxsize=None)
def small_rays():
    H, W, f = FERN_SMALL
    return O.make_rays(H, W, f, np.eye(4)[:3, :4])


@functools.lru_cache(maxsize=None)
def weights(kind="w1"):
    """W0: seed-0 default init (reference order).  W1: sigma-recalibrated on the golden probe rays.
    WD: default init with seed 16 -- the first seed whose last-sample sigma sign is mixed over the 1008x756 frame on BOTH
    nets (coarse 67 % / fine 25 % positive), i.e. a non-degenerate literal "random-init" set (most seeds
    give an all-or-nothing last-sample spike)."""
    if kind == "wd":
        return O.init_linear_like_reference(16)
    w0c, w0f = O.init_linear_like_reference(0)
    if kind == "w0":
        return w0c, w0f
    ro, rd = small_rays()
    probe = golden("chain_w1")["probe_index"]
    return O.recalibrate_sigma(w0c, ro[probe], rd[probe]), O.recalibrate_sigma(w0f, ro[probe], rd[probe])


def bf16r(x):
    return x.to(torch.bfloat16).to(torch.float32)


@torch.no_grad()
def emulate_bf16_forward(sd, pts, dirs_per_ray, S):
    """What mlp_tc.cu computes, restated with torch on the CPU: bf16 operands (PE, activations, weights),
    fp32 accumulation/bias/ReLU, fp32 sigma head on un-rounded h7, fp32 view-direction term, fp32 rgb1 head.
    pts [M,3]; dirs_per_ray [M/S,3].  Returns dict with per-layer fp32 accumulators (pre-bias) and outputs."""
    pe = O.embed(pts, 10).float()
    a_pe = bf16r(torch.cat([pe, torch.zeros_like(pe[:, :1])], -1))          # 64 columns
    W = lambda n: sd[n + ".weight"].float()
    B = lambda n: sd[n + ".bias"].float()
    accs = []
    w0 = bf16r(torch.cat([W("net.base_layers.0"), torch.zeros(256, 1)], -1))
    acc = a_pe @ w0.T
    accs.append(acc)
    h32 = torch.relu(acc + B("net.base_layers.0"))
    h = bf16r(h32)
    for i in range(1, 8):
        w = W("net.base_layers.%d" % i)
        if i == 5:
            w5 = bf16r(torch.cat([w[:, :63], torch.zeros(256, 1), w[:, 63:]], -1))
            acc = torch.cat([a_pe, h], -1) @ w5.T
        else:
            acc = h @ bf16r(w).T
        accs.append(acc)
        h32 = torch.relu(acc + B("net.base_layers.%d" % i))
        h = bf16r(h32)
    sigma = h32 @ W("net.sigma_layer").T + B("net.sigma_layer")
    acc = h @ bf16r(W("net.base_remap_layer")).T
    accs.append(acc)
    remap = bf16r(torch.relu(acc + B("net.base_remap_layer")))
    wr = W("net.rgb_layers.0")
    acc = remap @ bf16r(wr[:, :256]).T
    accs.append(acc)
    de = O.embed(dirs_per_ray, 4).float()
    dirbias = de @ wr[:, 256:].T + B("net.rgb_layers.0")                     # [rays,128] fp32
    f = torch.relu(acc + dirbias.repeat_interleave(S, dim=0))
    rgb = torch.sigmoid(f @ W("net.rgb_layers.1").T + B("net.rgb_layers.1"))
    return {"accs": accs, "rgb": rgb, "sigma": sigma.squeeze(-1)}


def knife_edge_mask(sigma, ts, rel=0.01, frac=0.002, tol=1e-2):
    """H1 protocol: a ray is knife-edge when the ORACLE's own composite (rgb of an all-ones colour field = acc)
    moves by more than the tolerance under a small sigma perturbation -- delta_last = 1e10 (utils.py:369) makes alpha_last a
    step function of the sign of the last sigma, so no finite-precision MLP can be held to 1e-2 there.  The perturbation is
    +-(1 % of |sigma| + 0.2 % of the field's sigma scale (its std)) per sample, all samples shifted the same way; the second
    term is what lets a near-zero last sigma change sign (a purely relative perturbation never would)."""
    rgbp = torch.ones(sigma.shape + (3,))
    base = O.alpha_composition(rgbp, sigma, ts)[3]
    d = rel * sigma.abs() + frac * sigma.std()
    lo = O.alpha_composition(rgbp, sigma - d, ts)[3]
    hi = O.alpha_composition(rgbp, sigma + d, ts)[3]
    return ((lo - base).abs() > tol) | ((hi - base).abs() > tol)


@functools.lru_cache(maxsize=None)
def fullsize_rays(n_per_frame=8192, seed=0):
    """n_per_frame rays spread over the 1008x756 identity frame and n_per_frame over spiral pose 17 (BASELINE configs 2 / 3)."""
    H, W, f = FERN_FULL
    rng = np.random.RandomState(seed)
    ros, rds = [], []
    for pose in (np.eye(4)[:3, :4], golden("rays")["spiral_poses"][17]):
        ro, rd = O.make_rays(H, W, f, pose)
        sel = np.sort(rng.choice(ro.shape[0], n_per_frame, replace=False))
        ros.append(ro[sel])
        rds.append(rd[sel])
    return np.concatenate(ros), np.concatenate(rds)


@functools.lru_cache(maxsize=None)
def fullsize_reference(kind, n_per_frame=8192):
    """the oracle's chain (rendering.py:27-51) on fullsize_rays, every intermediate kept"""
    ro, rd = fullsize_rays(n_per_frame)
    wc, wf = weights(kind)
    return O.render_chain(wc, wf, ro, rd, 0., 1., 64, 64, 4096, keep_intermediates=True)
