"""CPU: the N>1 host path (ray sharding + tile all-gather) on world_size=2/3 gloo process groups."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings, strategies as st

import tgtc_style_b200 as T


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10 ** 7), world=st.integers(1, 16))
def test_shard_ranges_partition(n, world):
    ranges = [T.shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1
    sizes = [e - b for b, e in ranges]
    assert max(sizes) - min(sizes) <= 1 and sizes == T.shard_sizes(n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(begin, end):
    """stand-in for NerfRenderer.render on a ray range: deterministic functions of the ray index."""
    idx = torch.arange(begin, end, dtype=torch.float32)
    return {"rgb": torch.stack([idx, idx * 2, idx * 3], -1), "depth": idx + 0.5, "acc": torch.ones_like(idx)}


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = T.render_frame_sharded(_fake_render, n_total, gather=True)
        ref = _fake_render(0, n_total)
        ok = all(torch.equal(full[k], ref[k]) for k in ref)
        local = T.render_frame_sharded(_fake_render, n_total, gather=False)
        b, e = T.shard_range(n_total, rank, world)
        ok = ok and local["rgb"].shape[0] == e - b
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1008 * 8), (2, 1001), (3, 1000)])
def test_sharded_render_gather_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _gather_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tg = T.TileGatherer(n, "cpu")
        ok = True
        handles = []
        for frame in range(3):          # three frames through two buffer pairs: the first pair is reused
            out = tg.outputs()
            assert out["rgb"].shape == (n, 3) and out["depth"].shape == (n,) and out["acc"].shape == (n,)
            src = _fake_render(frame * 1000 + rank * n, frame * 1000 + (rank + 1) * n)
            for k in out:
                out[k].copy_(src[k])     # what NerfRenderer.render(out=...) does: writes into the packed send buffer
            handles.append((frame, tg.gather_async()))
            f, h = handles[-1]
            full = h.wait()
            for r in range(world):
                ref = _fake_render(f * 1000 + r * n, f * 1000 + (r + 1) * n)
                ok = ok and all(torch.equal(full[k][r], ref[k]) for k in ref)
        tg.finish()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_packed_tile_gather_gloo():
    world, n = 2, 257
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(r, True) for r in range(world)]


# ------------------------------------------------------------------ data-parallel training step (host logic, gloo)
class _FakeRenderer:
    """CPU stand-in for NerfRenderer in NerfTrainer: the 'gradient' is a deterministic linear function of the rays,
    so sharding + all-reduce must reproduce the single-process result."""
    P = 595844

    def __init__(self):
        self.device = torch.device("cpu")
        self.set_calls = 0

    def grad_views(self, flat):
        from tgtc_style_b200.render import LAYER_NAMES, LAYER_SHAPES
        out = []
        for net in range(2):
            d, o = {}, net * self.P
            for name, (no, ni) in zip(LAYER_NAMES, LAYER_SHAPES):
                d[name + ".weight"] = flat[o:o + no * ni].view(no, ni); o += no * ni
                d[name + ".bias"] = flat[o:o + no]; o += no
            out.append(d)
        return tuple(out)

    def set_weights(self, coarse=None, fine=None):
        self.set_calls += 1

    def train_step(self, ro, rd, gt, n_total=None, grads=None, accumulate=False, **kw):
        basis = torch.linspace(-1, 1, 2 * self.P, dtype=torch.float64)
        s = (ro.double().sum() + 2 * rd.double().sum() + 3 * gt.double().sum()) / n_total
        g = (basis * s).float()
        if accumulate:
            grads += g
        else:
            grads.copy_(g)
        return {"loss": (gt.double() ** 2).sum().float() / (3 * n_total), "grads": grads}


def _sd():
    from tgtc_style_b200.render import LAYER_NAMES, LAYER_SHAPES
    g = torch.Generator().manual_seed(0)
    d = {}
    for name, (o, i) in zip(LAYER_NAMES, LAYER_SHAPES):
        d[name + ".weight"] = torch.randn(o, i, generator=g) * 0.1
        d[name + ".bias"] = torch.randn(o, generator=g) * 0.1
    return d


def _train_batch(n):
    g = torch.Generator().manual_seed(1)
    return torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g)


def _train_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = T.NerfTrainer(_FakeRenderer(), _sd(), _sd(), max_rays_per_pass=97)
        ro, rd, gt = _train_batch(n)
        tr.step(ro, rd, gt)
        flat = torch.cat([p.detach().flatten() for d in tr.params for p in d.values()])
        q.put((rank, flat[::997].clone().numpy(), tr.grads[::997].clone().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (2, 333), (2, 1)])     # n = 1 < world: rank 1 owns an EMPTY shard
def test_trainer_data_parallel_matches_single_process(world, n):
    single = T.NerfTrainer(_FakeRenderer(), _sd(), _sd(), max_rays_per_pass=97)
    ro, rd, gt = _train_batch(n)
    single.step(ro, rd, gt)
    ref_p = torch.cat([p.detach().flatten() for d in single.params for p in d.values()])[::997].numpy()
    ref_g = single.grads[::997].numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert len(res) == world
    for rank, p_s, g_s in res:
        np.testing.assert_allclose(g_s, ref_g, rtol=2e-5, atol=1e-7)      # summed shard gradients == full-batch gradients
        np.testing.assert_allclose(p_s, ref_p, rtol=1e-5, atol=1e-6)      # identical Adam step on every rank


# ------------------------------------------------------------------ Style_train data parallelism (host logic, gloo)
class _FakeStyleRenderer:
    """CPU stand-in for NerfRenderer in StyleTrainer: forward and backward are per-ray separable and linear in the upstream
    gradients, so equal shards + the trainer's all-reduces must reproduce the single-process iteration."""
    STYLE_C_SHAPES = T.NerfRenderer.STYLE_C_SHAPES
    STYLE_W_SHAPES = T.NerfRenderer.STYLE_W_SHAPES
    style_grad_views = T.NerfRenderer.style_grad_views

    def __init__(self):
        self.device = torch.device("cpu")
        self.P = sum(o * i + o for o, i in list(self.STYLE_C_SHAPES) + list(self.STYLE_W_SHAPES))
        self.basis = torch.linspace(-1, 1, self.P, dtype=torch.float64)

    def style_num_params(self):
        return self.P

    def style_train_workspace_bytes(self, n, n_samples=64, n_fine=64):
        return 16

    def set_style_weights(self, concat_style, style):
        pass

    def style_train_forward(self, ro, rd, lat, rand=None, noise_coarse=None, noise_fine=None, workspace=None, **kw):
        return {"rgb_coarse": torch.sigmoid(0.3 * ro + lat[:, :3]), "rgb_fine": torch.sigmoid(0.2 * rd + lat[:, 3:6]),
                "state": {"ro": ro, "rd": rd, "n": ro.shape[0]}}

    def style_train_backward(self, state, gc, gf, grads=None, accumulate=False):
        s = (gc.double() * state["ro"].double()).sum() + 2.0 * (gf.double() * state["rd"].double()).sum()
        g = (self.basis * s).float()
        if accumulate:
            grads += g
        else:
            grads.copy_(g)
        dl = torch.zeros(state["n"], 32)
        dl[:, :3], dl[:, 3:6] = gc, gf
        return {"grads": grads, "d_latents": dl}

    def adam_step(self, params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        exp_avg.mul_(betas[0]).add_(grads, alpha=1 - betas[0])
        exp_avg_sq.mul_(betas[1]).addcmul_(grads, grads, value=1 - betas[1])
        denom = (exp_avg_sq.sqrt() / (1 - betas[1] ** step) ** 0.5).add_(eps)
        params.addcdiv_(exp_avg, denom, value=-lr / (1 - betas[0] ** step))


def _style_sd(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    d = {}
    for i, (o, k) in enumerate(shapes):
        d["layers.%d.weight" % i] = torch.randn(o, k, generator=g) * 0.05
        d["layers.%d.bias" % i] = torch.randn(o, generator=g) * 0.05
    return d


def _style_iteration_inputs(n):
    g = torch.Generator().manual_seed(3)
    out = []
    for it in range(2):
        pair = []
        for origin in (False, True):
            b = {"rays_o": torch.randn(n, 3, generator=g), "rays_d": torch.randn(n, 3, generator=g), "rgb_gt": torch.rand(n, 3, generator=g),
                 "style_id": torch.randint(0, 2, (n,), generator=g), "frame_id": torch.randint(0, 4, (n,), generator=g)}
            if origin:
                b["rgb_origin"] = torch.rand(n, 3, generator=g)
            pair.append(b)
        out.append(pair)
    table = torch.randn(2, 4, 32, generator=g) * 0.5
    mu, logvar = torch.randn(2, 32, generator=g) * 0.3, torch.randn(2, 32, generator=g) * 0.2
    return out, table, mu, logvar


def _style_run(n, rank=0, world=1):
    its, table, mu, logvar = _style_iteration_inputs(n)
    r = _FakeStyleRenderer()
    lat = T.StyleLatents(table, mu, logvar, dataset_type="llff")
    tr = T.StyleTrainer(r, _style_sd(r.STYLE_C_SHAPES, 1), _style_sd(r.STYLE_W_SHAPES, 2), lat, frame_num=4)
    b, e = rank * n // world, (rank + 1) * n // world
    losses = None
    for pair in its:
        losses = tr.step(*({k: v[b:e] for k, v in d.items()} for d in pair))
    return (tr.grads[::991].clone().numpy(), tr.flat[::991].clone().numpy(), lat.latents.detach().clone().numpy(),
            {k: float(v) for k, v in losses.items()})


def _style_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank,) + _style_run(n, rank, world))
    finally:
        dist.destroy_process_group()


def test_style_trainer_data_parallel_matches_single_process():
    n, world = 64, 2
    g_ref, p_ref, tab_ref, l_ref = _style_run(n)
    assert l_ref["loss_coh"] > 0.0                      # the second iteration carries the coherence term
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_style_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, g_s, p_s, tab, losses in res:
        np.testing.assert_allclose(g_s, g_ref, rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(p_s, p_ref, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(tab, tab_ref, rtol=1e-4, atol=2e-6)
        for k in ("loss", "loss_rgb", "loss_logp", "loss_coh"):
            assert abs(losses[k] - l_ref[k]) <= 1e-4 * max(1.0, abs(l_ref[k])), (k, losses[k], l_ref[k])


def test_style_trainer_coherence_bookkeeping_matches_reference_loop():
    """StyleTrainer._coh_active against a transcription of the reference's `cnt` logic (train_tgtcs.py:347, :397-404, :449-459):
    which iterations compare the coherence batch with the previous one"""
    def reference_sequence(frame_num, iters):
        cnt, out = 0, []
        for _ in range(iters):
            coarse_term = False
            if cnt == frame_num:                  # :397-399
                pass
            else:
                if cnt != 0:                      # :400-401
                    coarse_term = True
            fine_term = False
            if cnt == frame_num:                  # :449-451
                cnt = 1
            else:
                if cnt != 0:                      # :453-456
                    fine_term = True
                cnt += 1                          # :458
            assert coarse_term == fine_term
            out.append(coarse_term)
        return out

    for frame_num in (1, 3, 5):
        its, table, mu, logvar = _style_iteration_inputs(8)
        r = _FakeStyleRenderer()
        tr = T.StyleTrainer(r, _style_sd(r.STYLE_C_SHAPES, 1), _style_sd(r.STYLE_W_SHAPES, 2), T.StyleLatents(table, mu, logvar), frame_num=frame_num)
        got = []
        for i in range(12):
            out = tr.step(*its[i % 2])
            got.append(float(out["loss_coh"]) > 0.0)
        assert got == reference_sequence(frame_num, 12), (frame_num, got)
