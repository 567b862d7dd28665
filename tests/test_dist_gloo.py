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
