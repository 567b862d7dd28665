"""Stand-alone HBM-roofline numbers of the byte-bound stage kernels at the full frame (762 048 rays): ray-gen + NDC (24 B/ray),
sample_pdf + sorted union (768 B/ray), compositing coarse / fine (1 556 / 3 092 B/ray) -- algorithmic bytes of SURVEY.md 8(d) over
the CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.    python tools/stage_bench.py > profiles/stage_bench_r2.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tgtc_style_b200 as T
from bench import synth_nerf_weights, H, W, FOCAL


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = 6545.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    r = T.NerfRenderer("cuda:0", mode="f16")
    r.set_weights(*synth_nerf_weights(0))
    K = np.array([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1]])
    n = H * W
    ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4])
    g = torch.Generator(device="cuda").manual_seed(0)
    res = {"rays": n, "hbm_peak_gbs": peak, "kernels": {}}

    def add(name, ms, bytes_per_ray):
        gbs = n * bytes_per_ray / (ms * 1e-3) / 1e9
        res["kernels"][name] = {"ms": ms, "bytes_per_ray": bytes_per_ray, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}

    add("raygen_kernel (K1)", timeit(lambda: r.raygen(H, W, K, np.eye(4)[:3, :4])), 24)
    # realistic coarse weights: a render's own
    out = r.render(ro, rd, 0., 1., extras=True, want_weights=False)
    wc, tsf = out["weights_coarse"], out["ts_fine"]
    ts_c = r.sample_uniform(ro[:1], rd[:1], 64, 0., 1., want_pts=False)[1][0].contiguous()
    add("sample_fine_kernel (K6+K7)", timeit(lambda: r.sample_fine(None, None, ts_c, wc, 64, want_pts=False)), 768)
    rs_c = torch.rand(n, 64, 4, device="cuda", generator=g)
    rs_f = torch.rand(n, 128, 4, device="cuda", generator=g)
    rs_c[..., 3] = (rs_c[..., 3] - 0.5) * 60
    rs_f[..., 3] = (rs_f[..., 3] - 0.5) * 60
    add("composite_kernel coarse (K5, stand-alone)", timeit(lambda: r.composite(t_values=ts_c, rgbsigma=rs_c)), 1556)
    add("composite_kernel fine (K5, stand-alone)", timeit(lambda: r.composite(t_values=tsf, rgbsigma=rs_f)), 3092)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
