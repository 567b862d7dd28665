"""CPU study (no GPU): which bf16 rounding of the tcgen05 MLP drives the composite error of the teacher-forced fine pass?
Emulates mlp_tc.cu's arithmetic with torch (bf16 operands, fp32 accumulate) under several operand-splitting variants and
reports, against the fp32 oracle, the per-sample sigma error and the composite |drgb| / |dacc| / |ddepth| statistics over
rays of the 1008x756 frame (identity pose + spiral pose 17).
    python tools/precision_study.py [n_rays]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import render_oracle as O

bf = lambda x: x.to(torch.bfloat16).to(torch.float32)


def split(x, terms):
    """x ~ sum of `terms` bf16 values (hi, lo, ...)"""
    out, r = [], x
    for _ in range(terms):
        h = bf(r)
        out.append(h)
        r = r - h
    return out


def mm(a, w, a_terms=1, w_terms=1):
    """sum over the hi/lo cross terms that a split GEMM would issue (drops lo x lo)"""
    A, W = split(a, a_terms), split(w, w_terms)
    acc = A[0] @ W[0].T
    if a_terms > 1:
        acc = acc + A[1] @ W[0].T
    if w_terms > 1:
        acc = acc + A[0] @ W[1].T
    return acc


@torch.no_grad()
def emulate(sd, pts, dirs_per_ray, S, pe_a=1, pe_w=1, hid_a=1, hid_w=1, layers_split=()):
    pe = O.embed(pts, 10).float()
    W = lambda n: sd[n + ".weight"].float()
    B = lambda n: sd[n + ".bias"].float()
    h32 = torch.relu(mm(pe, W("net.base_layers.0"), pe_a, pe_w) + B("net.base_layers.0"))
    for i in range(1, 8):
        w = W("net.base_layers.%d" % i)
        ha, hw = (2, 2) if i in layers_split else (hid_a, hid_w)
        if i == 5:
            acc = mm(pe, w[:, :63], pe_a, pe_w) + mm(h32, w[:, 63:], ha, hw)
        else:
            acc = mm(h32, w, ha, hw)
        h32 = torch.relu(acc + B("net.base_layers.%d" % i))
    sigma = h32 @ W("net.sigma_layer").T + B("net.sigma_layer")
    remap = torch.relu(mm(h32, W("net.base_remap_layer")) + B("net.base_remap_layer"))
    wr = W("net.rgb_layers.0")
    de = O.embed(dirs_per_ray, 4).float()
    dirbias = de @ wr[:, 256:].T + B("net.rgb_layers.0")
    f = torch.relu(mm(remap, wr[:, :256]) + dirbias.repeat_interleave(S, dim=0))
    rgb = torch.sigmoid(f @ W("net.rgb_layers.1").T + B("net.rgb_layers.1"))
    return rgb, sigma.squeeze(-1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    H, W_, f = 756, 1008, 815.13
    import helpers
    poses = [np.eye(4)[:3, :4]]
    g = np.random.RandomState(0)
    rays = []
    for p in poses:
        ro, rd = O.make_rays(H, W_, f, p)
        sel = g.choice(ro.shape[0], n, replace=False)
        rays.append((ro[sel], rd[sel]))
    ro = np.concatenate([r[0] for r in rays]); rd = np.concatenate([r[1] for r in rays])
    for kind in ("w1",):
        wc, wf = helpers.weights(kind)
        ref = O.render_chain(wc, wf, ro, rd, 0., 1., 64, 64, 4096, keep_intermediates=True)
        ts = ref["ts_fine"]; sig_ref = ref["sigma_fine"]; N = ts.shape[0]
        pts = ref["pts_fine"].reshape(-1, 3)
        flagged = helpers.knife_edge_mask(sig_ref, ts)
        scale = sig_ref.std().item()
        # alternative flag: +-1% of the sigma scale
        ones = torch.ones(sig_ref.shape + (3,))
        base = O.alpha_composition(ones, sig_ref, ts)[3]
        fl2 = torch.zeros_like(flagged)
        for s in (+0.01 * scale, -0.01 * scale):
            fl2 |= (O.alpha_composition(ones, sig_ref + s, ts)[3] - base).abs() > 1e-2
        print("%s: rays %d sigma scale %.2f flagged(rel 1%%) %.4f flagged(1%% of scale) %.4f" % (kind, N, scale, flagged.float().mean(), fl2.float().mean()))
        variants = {
            "all bf16 (today)": dict(),
            "PE A split": dict(pe_a=2),
            "PE A+W split": dict(pe_a=2, pe_w=2),
            "hidden A split": dict(hid_a=2),
            "hidden W split": dict(hid_w=2),
            "hidden A+W split": dict(hid_a=2, hid_w=2),
            "everything split": dict(pe_a=2, pe_w=2, hid_a=2, hid_w=2),
            "L7 A+W split": dict(layers_split=(7,)),
            "L6,L7 A+W split": dict(layers_split=(6, 7)),
            "L5-7 A+W split + PE": dict(layers_split=(5, 6, 7), pe_a=2, pe_w=2),
        }
        for name, kw in variants.items():
            rgb, sig = emulate(wf, pts, torch.from_numpy(rd), 128, **kw)
            rgb = rgb.reshape(N, 128, 3); sig = sig.reshape(N, 128)
            c = O.alpha_composition(rgb, sig, ts)
            e = torch.stack([(c[0] - ref["rgb"]).abs().max(-1)[0], (c[1] - ref["depth"]).abs(), (c[3] - ref["acc"]).abs()], 0).max(0)[0]
            ds = (sig - sig_ref).abs()
            print("  %-24s |dsigma| med %.3e max %.3e | composite err: mean %.2e p99 %.2e max %.2e  max(not flagged rel) %.2e n>1e-2 %d | max(not flagged scale) %.2e n>1e-2 %d"
                  % (name, ds.median(), ds.max(), e.mean(), torch.quantile(e, 0.99), e.max(), e[~flagged].max(), (e[~flagged] > 1e-2).sum(),
                     e[~fl2].max(), (e[~fl2] > 1e-2).sum()))


if __name__ == "__main__":
    main()
