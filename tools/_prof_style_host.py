import sys, os, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import tgtc_style_b200 as T
H, W, f = 756, 1008, 815.13
wc, wf = B.synth_nerf_weights(0); cs, ws = B.synth_style_weights(1)
K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])
dev = torch.device("cuda:0")
r = T.NerfRenderer(dev, mode="bf16"); r.set_weights(wc, wf)
ro_all, rd_all = r.raygen(H, W, K, np.eye(4)[:3, :4])
gen = torch.Generator().manual_seed(4)
table = torch.randn(4, 20, 32, generator=gen) * 0.5
mu, logvar = torch.randn(4, 32, generator=gen) * 0.3, torch.randn(4, 32, generator=gen) * 0.2
lat = T.StyleLatents(table.to(dev), mu.to(dev), logvar.to(dev))
tr = T.StyleTrainer(r, cs, ws, lat, frame_num=20)
batches = B._style_train_batches(ro_all, rd_all, 256, 4, 20, gen, dev)
for s in range(4): tr.step(*batches[s % 4])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for s in range(50): tr.step(*batches[s % 4])
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
