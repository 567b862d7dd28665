"""Device timing of the stylised render (render_style loop body).  Development aid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
from quick_bench import timeit

def main():
    ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 65536]
    H, W, f = 756, 1008, 815.13
    w0c, w0f = B.synth_nerf_weights(0)
    cs, ws = B.synth_style_weights(1)
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
    r = T.NerfRenderer("cuda:0", mode="bf16")
    r.set_weights(w0c, w0f)
    r.set_style_weights(cs, ws)
    lat = torch.randn(32, device="cuda")
    for n in ns:
        ro, rd = r.raygen(H, W, K, np.eye(4)[:3, :4], pix_begin=0, n=n)
        for chunk in (None, 4096):
            if chunk and n <= chunk:
                continue
            r.profile_enable(True)
            ms = timeit(lambda: r.render_style(ro, rd, lat, chunk=chunk), iters=3, warm=1)
            kinds = [r.profile_read_kind(k) for k in range(3)]
            r.profile_enable(False)
            fl = n * 192 * 2.0 * (593408 - 36224 - 384 + 335360 + 614752)
            print("style render n=%d chunk=%s: %.3f ms  %.3f Mrays/s  %.1f TFLOP/s | trunk %.2f ms %.0f TF, module1 %.2f ms %.0f TF, module2 %.2f ms %.0f TF" % (
                n, chunk, ms, n / ms / 1e3, fl / ms / 1e9, kinds[0][1] / 4, kinds[0][2] / kinds[0][1] / 1e9, kinds[1][1] / 4,
                kinds[1][2] / kinds[1][1] / 1e9, kinds[2][1] / 4, kinds[2][2] / kinds[2][1] / 1e9))

if __name__ == "__main__":
    main()
