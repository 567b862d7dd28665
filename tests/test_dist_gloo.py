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


@pytest.mark.parametrize("world,n", [(2, 1000), (2, 333)])
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
