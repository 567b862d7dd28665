"""TEST INFRASTRUCTURE ONLY -- CPU restatement of TGTC-Style's NeRF ray-render hot path.

This file is the *checker*.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product
(tgtc-style_b200/) never does: it fails loudly if its CUDA library is missing.

What it restates (all citations relative to /root/reference):
  dataset.py:33-42    get_rays_np           -> get_rays
  dataset.py:44-61    ndc_rays_np           -> ndc_rays
  utils.py:509-531    sampling_pts_uniform  -> sample_uniform
  models.py:24-60     Embedder              -> embed
  models.py:63-117    MLP_style             -> mlp_forward
  models.py:182-223   StyleNerf             -> nerf_forward
  utils.py:354-386    alpha_composition     -> alpha_composition
  utils.py:583-609    sample_pdf            -> sample_pdf
  utils.py:573-580    sampling_pts_fine_torch -> sample_fine
  rendering.py:27-51  cal_geometry loop body -> render_chain

The arithmetic of the reference lives in PyTorch (un-pinned in its
requirements.txt); the oracle therefore uses CPU torch fp32 ops for everything
whose exact bits do not matter, and *explicit* numpy restatements for the three
places where the reference's bits decide an integer result
(`torch.linspace`, `torch.sum(-1)` and `torch.cumsum` inside sample_pdf) so that
the oracle is deterministic on any host:
  * linspace : step=(end-start)/(steps-1); x_i = fma(step,i,start) for i<steps/2
               else fma(-step,steps-1-i,end)            (verified == torch here)
  * sum(-1)  : ATen's vectorized_inner_sum with 8-lane vectors and 4-way ILP
               (see row_sum_f32)                        (verified == torch here)
  * cumsum   : accumulate in fp64, round every prefix   (verified == torch here)

Parity pin: the reference has no tests or golden vectors, so
the oracle is pinned by executing the reference itself in the build container:
tests/test_oracle_vs_reference.py (live, when /root/reference exists) and
tests/golden/*.npz (vectors produced by oracle/make_golden.py from the imported
reference; checked everywhere).
"""
from collections import OrderedDict

import numpy as np
import torch

# ----------------------------------------------------------------------------
# explicit restatements of the three bit-deciding torch ops


def linspace_f32(start, end, steps):
    """torch.linspace(start, end, steps) for fp32 on CPU, bit for bit."""
    start = np.float32(start)
    end = np.float32(end)
    if steps == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((end - start) / np.float32(steps - 1))
    idx = np.arange(steps)
    # fp32*int products are exact in fp64, so fp64-then-round == fused multiply-add
    lo = (np.float64(start) + np.float64(step) * idx).astype(np.float32)
    hi = (np.float64(end) - np.float64(step) * (steps - 1 - idx)).astype(np.float32)
    return np.where(idx < steps // 2, lo, hi).astype(np.float32)


def row_sum_f32(a):
    """torch.sum(a, -1) for a contiguous fp32 [N, n] CPU tensor, bit for bit.

    ATen (SumKernel.cpp, vectorized_inner_sum -> row_sum -> multi_row_sum) walks
    each row as n//8 vectors of 8 lanes.  The first 4*(nvec//4) vectors feed four
    independent accumulators (vector j goes to accumulator j%4, in order); the
    left-over vectors are added to accumulator 0; accumulators 1,2,3 are then
    added to accumulator 0; finally a scalar starts from the sum of the n%8
    tail elements (in order) and adds the 8 lanes in order.  (The cascade levels
    of multi_row_sum only matter for nvec//4 > 16, i.e. n >= 544.)
    """
    a = np.ascontiguousarray(a, dtype=np.float32)
    rows, n = a.shape
    assert 8 <= n < 544, "scalar path (n<8) and cascade levels (n>=544) not restated"
    nvec = n // 8
    nilp = nvec // 4
    acc = [np.zeros((rows, 8), np.float32) for _ in range(4)]
    for i in range(nilp):
        for k in range(4):
            j = 4 * i + k
            acc[k] = acc[k] + a[:, 8 * j:8 * j + 8]
    for j in range(4 * nilp, nvec):
        acc[0] = acc[0] + a[:, 8 * j:8 * j + 8]
    for k in range(1, 4):
        acc[0] = acc[0] + acc[k]
    fin = np.zeros(rows, np.float32)
    for k in range(8 * nvec, n):
        fin = fin + a[:, k]
    for k in range(8):
        fin = fin + acc[0][:, k]
    return fin


def cumsum_f32(a):
    """torch.cumsum(a, -1) for fp32 on CPU: fp64 accumulator, every prefix rounded."""
    return np.cumsum(np.asarray(a, np.float32).astype(np.float64), axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------
# rays (dataset.py:33-61) -- NumPy fp64, exactly like the reference


def get_rays(H, W, K, c2w, pixel_alignment=False):
    """dataset.py:33-42.  Returns fp64 (rays_o, rays_d) of shape [H, W, 3]."""
    K = np.asarray(K, np.float64)
    c2w = np.asarray(c2w, np.float64)
    col, row = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    if pixel_alignment:
        col, row = col + 0.5, row + 0.5
    cam = np.stack([(col - K[0][2]) / K[0][0], -(row - K[1][2]) / K[1][1], -np.ones_like(col)], axis=-1)
    rays_d = np.sum(cam[..., np.newaxis, :] * c2w[:3, :3], axis=-1)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """dataset.py:44-61 (forward-facing NDC warp), same operation order."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return np.stack([o0, o1, o2], -1), np.stack([d0, d1, d2], -1)


def make_rays(H, W, focal, c2w, ndc=True, pixel_alignment=False):
    """The contract's ray source: fp64 ray-gen + NDC(near=1), then cast to fp32
    (dataset.py:92-96 K construction, :105-118).  Returns [H*W, 3] fp32 x2."""
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], np.float64)
    ro, rd = get_rays(H, W, K, c2w, pixel_alignment)
    if ndc:
        ro, rd = ndc_rays(H, W, K[0][0], 1., ro, rd)
    return (np.ascontiguousarray(ro.reshape(-1, 3), dtype=np.float32),
            np.ascontiguousarray(rd.reshape(-1, 3), dtype=np.float32))


# ----------------------------------------------------------------------------
# sampling (utils.py:509-531)


def sample_uniform(rays_o, rays_d, n_samples=64, near=0., far=1.05, rand=None):
    """utils.py:509-531 with harmony=False.  `rand` ([N,S] in [0,1)) replays the
    perturb=True branch (utils.py:518-524); None means perturb=False."""
    rays_o = torch.as_tensor(rays_o)
    rays_d = torch.as_tensor(rays_d)
    n = rays_o.shape[0]
    ts = torch.from_numpy(linspace_f32(0., 1., n_samples)).unsqueeze(0).expand(n, n_samples)
    ts = ts * (far - near) + near
    if rand is not None:
        mid = (ts[..., 1:] + ts[..., :-1]) / 2
        upper = torch.cat([mid, ts[..., -1:]], -1)
        lower = torch.cat([ts[..., :1], mid], -1)
        ts = lower + (upper - lower) * torch.as_tensor(rand)
    pts = rays_o.unsqueeze(1) + ts.unsqueeze(-1) * rays_d.unsqueeze(1)
    return pts, ts


# ----------------------------------------------------------------------------
# network (models.py)

LAYER_NAMES = (["net.base_layers.%d" % i for i in range(8)]
               + ["net.sigma_layer", "net.base_remap_layer", "net.rgb_layers.0", "net.rgb_layers.1"])
LAYER_SHAPES = ([(256, 63)] + [(256, 256)] * 4 + [(256, 319)] + [(256, 256)] * 2
                + [(1, 256), (256, 256), (128, 283), (3, 128)])


def embed(x, n_freqs):
    """models.py:46-60: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] on the last dim."""
    out = [x]
    for k in range(n_freqs):
        f = float(2. ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def init_linear_like_reference(seed=0):
    """Weight set W0: torch.manual_seed(seed); StyleNerf('coarse'); StyleNerf('fine')
    (train_tgtcs.py:25-37 order).  MLP_style.__init__ (models.py:75-91) constructs
    nn.Linear modules in the order base_layers[0..7], sigma_layer,
    base_remap_layer, rgb_layers[0..1]; building the same nn.Linear sequence
    consumes the global RNG identically, so the tensors are bit-identical to the
    reference's (checked in tests/test_oracle_vs_reference.py)."""
    torch.manual_seed(seed)
    nets = []
    for _ in range(2):
        sd = OrderedDict()
        for name, (out_f, in_f) in zip(LAYER_NAMES, LAYER_SHAPES):
            lin = torch.nn.Linear(in_f, out_f)
            sd[name + ".weight"] = lin.weight.detach().clone()
            sd[name + ".bias"] = lin.bias.detach().clone()
        nets.append(sd)
    return nets[0], nets[1]


def mlp_forward(sd, pts_emb, dirs_emb):
    """models.py:95-117 (relu net, use_viewdir, skips=[4], sigma_mul=0)."""
    lin = torch.nn.functional.linear
    h = torch.relu(lin(pts_emb, sd["net.base_layers.0.weight"], sd["net.base_layers.0.bias"]))
    for i in range(7):
        if i == 4:
            h = torch.cat((pts_emb, h), dim=-1)
        h = torch.relu(lin(h, sd["net.base_layers.%d.weight" % (i + 1)], sd["net.base_layers.%d.bias" % (i + 1)]))
    sigma = lin(h, sd["net.sigma_layer.weight"], sd["net.sigma_layer.bias"])
    remap = torch.relu(lin(h, sd["net.base_remap_layer.weight"], sd["net.base_remap_layer.bias"]))
    f = torch.relu(lin(torch.cat((remap, dirs_emb), dim=-1), sd["net.rgb_layers.0.weight"], sd["net.rgb_layers.0.bias"]))
    rgb = torch.sigmoid(lin(f, sd["net.rgb_layers.1.weight"], sd["net.rgb_layers.1.bias"]))
    return OrderedDict([("rgb", rgb), ("base_remap", remap), ("pts", pts_emb), ("sigma", sigma.squeeze(-1))])


def nerf_forward(sd, pts, dirs):
    """models.py:216-223: embed (L=10 / L=4), cast fp32, MLP, add ret['dirs']."""
    pe = embed(pts, 10).to(torch.float32)
    de = embed(dirs, 4).to(torch.float32)
    ret = mlp_forward(sd, pe, de)
    ret["dirs"] = de
    return ret


def nerf_forward_chunked(sd, pts, dirs, chunk):
    """utils.py:435-456 batchify: chunk dim 0 (rays), cat every returned key."""
    outs = {}
    for i in range(0, pts.shape[0], chunk):
        r = nerf_forward(sd, pts[i:i + chunk], dirs[i:i + chunk])
        for k, v in r.items():
            outs.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in outs.items()}


def recalibrate_sigma(sd, rays_o, rays_d, gain=30.0, shift=0.0, near=0., far=1.):
    """Weight set W1: standardise sigma over a fixed probe
    batch (the first 256 rays given, 64 coarse samples) and rescale the sigma
    head so that sigma ~ N(shift, gain^2) on it.  Pure test-input synthesis."""
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    ro = torch.as_tensor(rays_o[:256])
    rd = torch.as_tensor(rays_d[:256])
    pts, _ = sample_uniform(ro, rd, 64, near, far)
    with torch.no_grad():
        sig = nerf_forward(sd, pts, rd.unsqueeze(1).expand(pts.shape))["sigma"]
    mu, s = sig.mean(), sig.std()
    sd["net.sigma_layer.weight"] = sd["net.sigma_layer.weight"] * (gain / s)
    sd["net.sigma_layer.bias"] = (sd["net.sigma_layer.bias"] - mu) * (gain / s) + shift
    return sd


# ----------------------------------------------------------------------------
# compositing (utils.py:354-386)


def alpha_composition(pts_rgb, pts_sigma, t_values, noise=None, white_bkgd=False):
    """utils.py:354-386.  `noise` replays `randn*sigma_noise_std` (utils.py:372-374).
    Returns (rgb_exp, t_exp, weights, acc_map); the reference computes acc_map
    (utils.py:382) and drops it."""
    delta = t_values[..., 1:] - t_values[..., :-1]
    delta = torch.cat([delta, torch.full_like(delta[..., :1], 1e10)], -1)
    sig = pts_sigma if noise is None else pts_sigma + noise
    alpha = 1. - torch.exp(-torch.relu(torch.relu(sig)) * delta)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_exp = torch.sum(weights[..., None] * pts_rgb, -2)
    t_exp = torch.sum(weights * t_values, -1)
    acc = torch.sum(weights, -1)
    if white_bkgd:
        rgb_exp = rgb_exp + (1. - acc[..., None])
    return rgb_exp, t_exp, weights, acc


# ----------------------------------------------------------------------------
# hierarchical resampling (utils.py:573-609)


def sample_pdf(bins, weights, n_samples):
    """utils.py:583-609 with det=True.  Returns (samples [N,n], inds [N,n] int64,
    cdf [N,nb]).  sum / cumsum / linspace use the explicit restatements above."""
    bins = np.ascontiguousarray(torch.as_tensor(bins).numpy(), dtype=np.float32)
    w = np.ascontiguousarray(torch.as_tensor(weights).numpy(), dtype=np.float32) + np.float32(1e-5)
    pdf = (w / row_sum_f32(w)[:, None]).astype(np.float32)
    cdf = np.concatenate([np.zeros_like(pdf[:, :1]), cumsum_f32(pdf)], -1)
    nb = cdf.shape[-1]
    u = linspace_f32(0., 1., n_samples)
    # searchsorted(cdf, u, right=True): number of cdf entries <= u
    inds = (cdf[:, None, :] <= u[None, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(0, inds - 1)
    above = np.minimum(nb - 1, inds)
    cdf_b = np.take_along_axis(cdf, below, 1)
    cdf_a = np.take_along_axis(cdf, above, 1)
    bins_b = np.take_along_axis(bins, below, 1)
    bins_a = np.take_along_axis(bins, above, 1)
    denom = (cdf_a - cdf_b).astype(np.float32)
    denom = np.where(denom < np.float32(1e-5), np.float32(1.), denom)
    t = ((u[None, :] - cdf_b) / denom).astype(np.float32)
    samples = (bins_b + (t * (bins_a - bins_b)).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(samples), torch.from_numpy(inds), torch.from_numpy(cdf)


def sample_fine(rays_o, rays_d, ts, weights, n_fine=64, return_aux=False):
    """utils.py:573-580: midpoints, sample_pdf on weights[1:-1], sort the union."""
    ts = torch.as_tensor(ts)
    weights = torch.as_tensor(weights)
    ts_mid = 0.5 * (ts[..., 1:] + ts[..., :-1])
    t_samples, inds, cdf = sample_pdf(ts_mid, weights[..., 1:-1], n_fine)
    t_vals = torch.sort(torch.cat([ts, t_samples], -1), -1)[0]
    pts = torch.as_tensor(rays_o).unsqueeze(-2) + torch.as_tensor(rays_d).unsqueeze(-2) * t_vals.unsqueeze(-1)
    if return_aux:
        return pts, t_vals, t_samples, inds, cdf
    return pts, t_vals


# ----------------------------------------------------------------------------
# the chain (rendering.py:27-51)


@torch.no_grad()
def render_chain(sd_coarse, sd_fine, rays_o, rays_d, near=0., far=1., n_samples=64, n_fine=64,
                 chunk=1024, keep_intermediates=False):
    """rendering.py:27-51 for one batch of rays (fp32), perturb off, noise 0.
    Returns the north-star surface {rgb, depth, acc, weights} (+ coarse outputs,
    ts_fine and, when asked, every per-sample intermediate)."""
    ro = torch.as_tensor(rays_o, dtype=torch.float32)
    rd = torch.as_tensor(rays_d, dtype=torch.float32)
    n = ro.shape[0]
    pts, ts = sample_uniform(ro, rd, n_samples, near, far)
    ret = nerf_forward_chunked(sd_coarse, pts, rd.unsqueeze(1).expand(n, n_samples, 3), chunk)
    rgb_c, t_c, w_c, acc_c = alpha_composition(ret["rgb"], ret["sigma"], ts)
    out = {"rgb_coarse": rgb_c, "depth_coarse": t_c, "acc_coarse": acc_c, "weights_coarse": w_c, "ts": ts.contiguous()}
    if keep_intermediates:
        out.update(pts_coarse=pts, rgb_pts_coarse=ret["rgb"], sigma_coarse=ret["sigma"])
    pts_f, ts_f, t_samples, inds, cdf = sample_fine(ro, rd, ts, w_c, n_fine, return_aux=True)
    ret_f = nerf_forward_chunked(sd_fine, pts_f, rd.unsqueeze(1).expand(n, n_samples + n_fine, 3), chunk)
    rgb_f, t_f, w_f, acc_f = alpha_composition(ret_f["rgb"], ret_f["sigma"], ts_f)
    out.update(rgb=rgb_f, depth=t_f, acc=acc_f, weights=w_f, ts_fine=ts_f, pdf_inds=inds)
    if keep_intermediates:
        out.update(pts_fine=pts_f, rgb_pts_fine=ret_f["rgb"], sigma_fine=ret_f["sigma"], t_samples=t_samples, cdf=cdf)
    return out


def train_step_reference(sd_coarse, sd_fine, rays_o, rays_d, rgb_gt, near=0., far=1., n_samples=64, n_fine=64, rand=None,
                         noise_coarse=None, noise_fine=None):
    """train_tgtcs.py:228-255 (Origin_train's step) with perturb=0 and sigma_noise_std=0, through torch.autograd:
    loss = mse(rgb_gt, rgb_coarse) + mse(rgb_gt, rgb_fine)  (utils.py:460).  No gradient flows through the
    resampling (utils.py:576-579).  Returns (loss, grads_coarse, grads_fine, rgb_coarse, rgb_fine, ts_fine)."""
    ro = torch.as_tensor(rays_o, dtype=torch.float32)
    rd = torch.as_tensor(rays_d, dtype=torch.float32)
    gt = torch.as_tensor(rgb_gt, dtype=torch.float32)
    n = ro.shape[0]
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in sd_coarse.items()}
    pf = {k: v.detach().clone().requires_grad_(True) for k, v in sd_fine.items()}
    pts, ts = sample_uniform(ro, rd, n_samples, near, far, rand=rand)
    ret = nerf_forward(pc, pts, rd.unsqueeze(1).expand(n, n_samples, 3))
    rgb_c, _, w_c, _ = alpha_composition(ret["rgb"], ret["sigma"], ts, noise=noise_coarse)
    pts_f, ts_f = sample_fine(ro, rd, ts, w_c.detach(), n_fine)
    ret_f = nerf_forward(pf, pts_f.detach(), rd.unsqueeze(1).expand(n, n_samples + n_fine, 3))
    rgb_f = alpha_composition(ret_f["rgb"], ret_f["sigma"], ts_f, noise=noise_fine)[0]
    loss = torch.mean((rgb_c - gt) ** 2) + torch.mean((rgb_f - gt) ** 2)
    loss.backward()
    return (loss.detach(), {k: v.grad for k, v in pc.items()}, {k: v.grad for k, v in pf.items()}, rgb_c.detach(), rgb_f.detach(),
            ts_f.detach())


# ----------------------------------------------------------------------------
# per-ray style head: models.StyleMLP_before_concat (models.py:120-147),
# models.StyleMLP_Wild_multilayers (models.py:149-180), called as in render_style (rendering.py:118-178)

STYLE_C_SHAPES = [(256, 95), (256, 288), (256, 288), (256, 288), (256, 351)]                      # "layers.{i}" of the concat module
STYLE_W_SHAPES = [(256, 607), (256, 288), (256, 288), (256, 288), (256, 351), (256, 288), (256, 288), (3, 288)]


def init_style_like_reference(seed=1):
    """torch.manual_seed(seed); StyleMLP_before_concat(args); StyleMLP_Wild_multilayers(args) with style_D=8, vae_latent=32,
    netwidth=256, embed_freq_coor=10 (train_tgtcs.py:41-52 order): the same nn.Linear sequence, so bit-identical tensors
    (checked against the imported reference in tests/test_oracle_golden.py)."""
    torch.manual_seed(seed)
    nets = []
    for shapes in (STYLE_C_SHAPES, STYLE_W_SHAPES):
        sd = OrderedDict()
        for i, (out_f, in_f) in enumerate(shapes):
            lin = torch.nn.Linear(in_f, out_f)
            sd["layers.%d.weight" % i] = lin.weight.detach().clone()
            sd["layers.%d.bias" % i] = lin.bias.detach().clone()
        nets.append(sd)
    return nets[0], nets[1]


def style_concat_forward(sd, x, latent):
    """StyleMLP_before_concat.forward (models.py:137-147): x = embedded pts [...,63], latent [...,32]."""
    lin = torch.nn.functional.linear
    h = x
    for i in range(5):
        h = torch.cat([h, latent], dim=-1)
        if i == 4:
            h = torch.cat([h, x], dim=-1)
        h = torch.relu(lin(h, sd["layers.%d.weight" % i], sd["layers.%d.bias" % i]))
    return h


def style_wild_forward(sd, x, concated, latent):
    """StyleMLP_Wild_multilayers.forward (models.py:165-180): concated = cat(base_remap, concat_features) [...,512]."""
    lin = torch.nn.functional.linear
    h = torch.cat([concated, x], dim=-1)
    for i in range(7):
        h = torch.cat([h, latent], dim=-1)
        if i == 4:
            h = torch.cat([h, x], dim=-1)
        h = torch.relu(lin(h, sd["layers.%d.weight" % i], sd["layers.%d.bias" % i]))
    h = torch.cat([h, latent], dim=-1)
    return torch.sigmoid(lin(h, sd["layers.7.weight"], sd["layers.7.bias"]))


@torch.no_grad()
def render_style_chain(sd_coarse, sd_fine, sd_concat, sd_wild, rays_o, rays_d, latents, near=0., far=1., n_samples=64, n_fine=64,
                       keep_intermediates=False):
    """The loop body of render_style (rendering.py:118-178) for one batch of rays with perturb=False: NeRF trunk ->
    base_remap, sigma, embedded pts; style module 1 on (pts_embed, latents); style module 2 on (pts_embed,
    cat(base_remap, concat_features), mean-over-latent-dim broadcast) -> stylised rgb; compositing with the NeRF sigma;
    resampling; the same on the fine net.  latents: [N,32] (what latents_model_1 returns, rendering.py:125)."""
    ro = torch.as_tensor(rays_o, dtype=torch.float32)
    rd = torch.as_tensor(rays_d, dtype=torch.float32)
    lat = torch.as_tensor(latents, dtype=torch.float32)
    n = ro.shape[0]
    lat2 = torch.mean(lat, dim=1, keepdim=True)          # rendering.py:126 (mean over the LATENT dim)
    out = {}

    def one_pass(sd_nerf, pts, S):
        ret = nerf_forward(sd_nerf, pts, rd.unsqueeze(1).expand(n, S, 3))
        l1 = lat.unsqueeze(1).expand(n, S, lat.shape[-1])
        cf = style_concat_forward(sd_concat, ret["pts"], l1)
        concated = torch.cat((ret["base_remap"], cf), dim=-1)
        l2 = lat2.unsqueeze(2).expand(n, S, lat.shape[-1])
        rgb_s = style_wild_forward(sd_wild, ret["pts"], concated, l2)
        return ret, cf, rgb_s

    pts, ts = sample_uniform(ro, rd, n_samples, near, far)
    ret, cf, rgb_s = one_pass(sd_coarse, pts, n_samples)
    rgb_c, t_c, w_c, acc_c = alpha_composition(rgb_s, ret["sigma"], ts)
    out.update(rgb_coarse=rgb_c, depth_coarse=t_c, acc_coarse=acc_c, weights_coarse=w_c)
    if keep_intermediates:
        out.update(concat_features_coarse=cf, rgb_pts_coarse=rgb_s, sigma_coarse=ret["sigma"])
    pts_f, ts_f = sample_fine(ro, rd, ts, w_c, n_fine)
    ret_f, cf_f, rgb_sf = one_pass(sd_fine, pts_f, n_samples + n_fine)
    rgb_f, t_f, w_f, acc_f = alpha_composition(rgb_sf, ret_f["sigma"], ts_f)
    out.update(rgb=rgb_f, depth=t_f, acc=acc_f, weights=w_f, ts_fine=ts_f)
    if keep_intermediates:
        out.update(concat_features_fine=cf_f, rgb_pts_fine=rgb_sf, sigma_fine=ret_f["sigma"])
    return out


def _style_forward_diff(sd_coarse, sd_fine, pc, pw, ro, rd, lat, near, far, n_samples, n_fine, rand):
    """Differentiable (w.r.t. pc, pw, lat) forward of one Style_train batch; NeRF nets constant, resampling detached."""
    n = ro.shape[0]
    lat2 = torch.mean(lat, dim=1, keepdim=True)          # train_tgtcs.py:410

    def one_pass(sd_nerf, pts, S):
        with torch.no_grad():
            ret = nerf_forward(sd_nerf, pts, rd.unsqueeze(1).expand(n, S, 3))
        l1 = lat.unsqueeze(1).expand(n, S, lat.shape[-1])
        cf = style_concat_forward(pc, ret["pts"], l1)
        concated = torch.cat((ret["base_remap"], cf), dim=-1)
        l2 = lat2.unsqueeze(2).expand(n, S, lat.shape[-1])
        return ret, style_wild_forward(pw, ret["pts"], concated, l2)

    pts, ts = sample_uniform(ro, rd, n_samples, near, far, rand=rand)
    ret, rgb_s = one_pass(sd_coarse, pts, n_samples)
    rgb_c, _, w_c, _ = alpha_composition(rgb_s, ret["sigma"], ts)
    pts_f, ts_f = sample_fine(ro, rd, ts, w_c.detach(), n_fine)
    ret_f, rgb_sf = one_pass(sd_fine, pts_f.detach(), n_samples + n_fine)
    rgb_f = alpha_composition(rgb_sf, ret_f["sigma"], ts_f)[0]
    return rgb_c, rgb_f


def style_train_forward_backward(sd_coarse, sd_fine, sd_concat, sd_wild, rays_o, rays_d, latents, g_rgb_coarse, g_rgb_fine, near=0., far=1.,
                                 n_samples=64, n_fine=64, rand=None):
    """One batch of Style_train (train_tgtcs.py:354-483) through torch.autograd: the forward of render_style_chain with
    per-ray latents [N,32] (latents_model_1 output, train_tgtcs.py:409) and stratified positions from `rand`
    (perturb=True, train_tgtcs.py:362), then the vector-Jacobian product of (rgb_coarse, rgb_fine) with the given
    upstream gradients -- i.e. what loss.backward() sends into the two style modules and the latents for any loss built
    on the two rgb maps (train_tgtcs.py:425, :480-483).  The NeRF nets are constants (style_optimizer does not hold
    them, train_tgtcs.py:54); no gradient flows through the resampling (utils.py:576-579).
    Returns (rgb_coarse, rgb_fine, grads_concat, grads_wild, d_latents)."""
    ro = torch.as_tensor(rays_o, dtype=torch.float32)
    rd = torch.as_tensor(rays_d, dtype=torch.float32)
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in sd_concat.items()}
    pw = {k: v.detach().clone().requires_grad_(True) for k, v in sd_wild.items()}
    lat = torch.as_tensor(latents, dtype=torch.float32).detach().clone().requires_grad_(True)
    rgb_c, rgb_f = _style_forward_diff(sd_coarse, sd_fine, pc, pw, ro, rd, lat, near, far, n_samples, n_fine, rand)
    obj = (rgb_c * torch.as_tensor(g_rgb_coarse, dtype=torch.float32)).sum() + (rgb_f * torch.as_tensor(g_rgb_fine, dtype=torch.float32)).sum()
    obj.backward()
    return (rgb_c.detach(), rgb_f.detach(), {k: v.grad for k, v in pc.items()}, {k: v.grad for k, v in pw.items()}, lat.grad)


def _cos_rows(a, b):
    """VGGNet.cosine_similarity (VGGNet.py:204-210)."""
    an = a / (torch.norm(a, dim=1, keepdim=True) + 1e-8)
    bn = b / (torch.norm(b, dim=1, keepdim=True) + 1e-8)
    return torch.sum(an * bn, dim=1)


def style_train_step_reference(sd_coarse, sd_fine, sd_concat, sd_wild, lat_table, mu, logvar, batch, coh_batch, prev, frame_num,
                               rgb_loss_lambda=1.0, logp_lambda=0.1, loss_coh_lambda=1e2, dataset_type="llff", near=0., far=1.):
    """Loss and gradients of one Style_train iteration (train_tgtcs.py:354-495) with the previous loss_coh batch's maps
    `prev` = (x, y, x_origin) as constants (or None: no coherence term).  batch / coh_batch: dicts with rays_o, rays_d,
    rgb_gt, style_id, frame_id, rand (+ rgb_origin).  lat_table [style_num, frame_num, 32]; forward and minus_logp of
    models.StyleLatents_variational (models.py:490-506, :531-537) with sigma_scale = 1.
    Returns (losses dict, grads_concat, grads_wild, d lat_table)."""
    pc = {k: v.detach().clone().requires_grad_(True) for k, v in sd_concat.items()}
    pw = {k: v.detach().clone().requires_grad_(True) for k, v in sd_wild.items()}
    tab = torch.as_tensor(lat_table, dtype=torch.float32).detach().clone().requires_grad_(True)

    def lat_of(sid, fid):
        t = tab.reshape(-1, tab.shape[-1])
        if dataset_type == "llff":
            t = t.repeat((7, 1))
        m = mu[sid]
        return m + 1.0 * (t[sid * frame_num + fid] - m)

    def fwd(b):
        ro = torch.as_tensor(b["rays_o"], dtype=torch.float32)
        rd = torch.as_tensor(b["rays_d"], dtype=torch.float32)
        return _style_forward_diff(sd_coarse, sd_fine, pc, pw, ro, rd, lat_of(b["style_id"], b["frame_id"]), near, far, 64, 64, b["rand"])

    rgb_c, rgb_f = fwd(batch)
    gt = batch["rgb_gt"]
    loss_rgb = rgb_loss_lambda * (torch.mean((rgb_c - gt) ** 2) + torch.mean((rgb_f - gt) ** 2))
    lat = lat_of(batch["style_id"], batch["frame_id"])
    loss_logp = logp_lambda * torch.sum((lat - mu[batch["style_id"]]) ** 2 / (torch.exp(0.5 * logvar[batch["style_id"]]) + 1e-3), -1).mean()
    loss_coh = torch.zeros(())
    if coh_batch is not None and prev is not None:
        c2, f2 = fwd(coh_batch)
        x, y, x_org = prev
        org2 = coh_batch["rgb_origin"]
        l2 = lambda v: torch.sqrt(torch.sum(v ** 2) + 1e-8)          # utils.py:459
        # train_tgtcs.py:401 / :456 -- the fine term compares with x_origin AFTER line 403 replaced it by this batch's originals
        loss_coh = l2(_cos_rows(c2, x) - _cos_rows(org2, x_org)) + l2(_cos_rows(f2, y) - _cos_rows(org2, org2))
    loss = loss_rgb + loss_logp + loss_coh_lambda * loss_coh         # loss_for_style (train_tgtcs.py:482): style_optimizer's objective
    loss.backward(retain_graph=True)
    # the latent table is optimised on loss = loss_rgb + loss_logp only: latents_model_1.optimize(loss) zeroes its gradient and
    # backpropagates without the coherence term (train_tgtcs.py:481, :495; models.py:544-549)
    grad_tab = torch.autograd.grad(loss_rgb + loss_logp, tab)[0]
    losses = {"loss": loss.detach(), "loss_rgb": loss_rgb.detach(), "loss_logp": loss_logp.detach(), "loss_coh": loss_coh.detach()}
    return losses, {k: v.grad for k, v in pc.items()}, {k: v.grad for k, v in pw.items()}, grad_tab
