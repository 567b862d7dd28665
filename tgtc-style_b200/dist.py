"""Multi-GPU ray sharding.

Rays are independent through the whole chain, so a frame is partitioned into
contiguous ray-index ranges, one per rank (one process per GPU), with NO
collective on the data path.  The only exchange is the final all-gather of the
rendered tiles {rgb, depth, acc} (20 B/ray).  Works with any torch.distributed
backend: nccl on the GPU box, gloo in the CPU tests (which exercise the
partition / gather logic with a stand-in render function).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced partition of [0, n): the first n % world ranks get one extra ray."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_tiles(local, n_total, group=None):
    """all-gather of per-rank tiles.  local: dict name -> [n_local, ...] tensor (same names/trailing dims on
    every rank).  Returns dict name -> [n_total, ...] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(local)
    sizes = shard_sizes(n_total, world)
    big = max(sizes)
    out = {}
    for name in sorted(local):
        t = local[name].contiguous()
        trail = tuple(t.shape[1:])
        if len(set(sizes)) == 1:
            full = torch.empty((n_total,) + trail, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(full, t, group=group)
        else:
            # uneven shards: pad every tile to the largest one, gather, drop the padding
            pad = torch.zeros((big,) + trail, dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            buf = torch.empty((world * big,) + trail, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, pad, group=group)
            full = torch.cat([buf[r * big:r * big + sizes[r]] for r in range(world)], 0)
        out[name] = full
    return out


class TileGatherer:
    """The render path's only exchange, off the critical path: every rank renders its tile {rgb, depth, acc}
    straight into ONE packed send buffer ([rgb 3n | depth n | acc n] floats -- the render outputs are views of it, nothing is
    copied), and ONE all-gather per frame runs on a side stream while the compute stream already renders the next frame.
    Two buffer pairs alternate; a pair is reused only after its gather has completed.

        g = TileGatherer(n_local, device)
        out = g.outputs()              # {"rgb": [n,3], "depth": [n], "acc": [n]} views of the current send buffer
        renderer.render(..., out=out)
        h = g.gather_async()           # returns at once; h.wait() -> {"rgb": [world,n,3], "depth": [world,n], "acc": [world,n]}
    """

    def __init__(self, n_local, device, group=None, depth=2):
        self.n, self.group = int(n_local), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dev = torch.device(device)
        self.send = [torch.empty(5 * self.n, dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.recv = [torch.empty(self.world * 5 * self.n, dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.cur = 0
        cuda = self.dev.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.dev) if cuda else None
        self.done = [None] * depth            # event: the gather out of pair i has completed

    def outputs(self):
        """views of the CURRENT send buffer; waits (on the compute stream, not the host) until the pair's previous gather is done"""
        i = self.cur
        if self.done[i] is not None:
            torch.cuda.current_stream(self.dev).wait_event(self.done[i])
        b, n = self.send[i], self.n
        return {"rgb": b[:3 * n].view(n, 3), "depth": b[3 * n:4 * n], "acc": b[4 * n:5 * n]}

    class Handle:
        def __init__(self, owner, i):
            self.o, self.i = owner, i

        def wait(self):
            """-> the gathered tiles of every rank (views of the receive buffer, valid until the pair is reused)"""
            o, i, n, w = self.o, self.i, self.o.n, self.o.world
            if o.done[i] is not None:
                torch.cuda.current_stream(o.dev).wait_event(o.done[i])
            r = o.recv[i].view(w, 5 * n)
            return {"rgb": r[:, :3 * n].reshape(w, n, 3), "depth": r[:, 3 * n:4 * n], "acc": r[:, 4 * n:5 * n]}

    def gather_async(self):
        i = self.cur
        self.cur = (self.cur + 1) % len(self.send)
        if self.world == 1:
            self.recv[i].copy_(self.send[i])
            return TileGatherer.Handle(self, i)
        if self.comm is None:                       # CPU tensors (gloo tests): same protocol, synchronous
            dist.all_gather_into_tensor(self.recv[i], self.send[i], group=self.group)
            return TileGatherer.Handle(self, i)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            dist.all_gather_into_tensor(self.recv[i], self.send[i], group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        self.done[i] = ev
        return TileGatherer.Handle(self, i)

    def finish(self):
        """compute stream waits for every gather in flight (end of a path / before timing stops)"""
        if self.comm is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.comm)


def render_frame_sharded(render_fn, n_total, group=None, gather=True):
    """render_fn(begin, end) -> dict of per-ray tensors for rays [begin, end).  Each rank renders its
    contiguous range; with gather=True the tiles are all-gathered so every rank holds the frame."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    b, e = shard_range(n_total, rank, world)
    local = render_fn(b, e)
    if not gather:
        return local
    return gather_tiles(local, n_total, group)


def render_path_sharded(renderer, H, W, K, poses, group=None, split="rows", latents=None, **render_kw):
    """The spiral / validation path of the reference (render_style's frame loop, rendering.py:109-217, over the
    load_llff.render_path_spiral poses) on one or more GPUs.  Rays are generated on the device per pose
    (NerfRenderer.render_frame), frames stay on the device.
      split="rows"   every frame is split into contiguous ray ranges, one per rank, and all-gathered (lowest latency per frame)
      split="frames" whole frames go round-robin to ranks; the tiles of `world` frames are exchanged with ONE all-gather
    latents: None = plain NeRF frames (cal_geometry); a [32] tensor or a per-frame sequence of [32] tensors = stylised frames
    (render_valid_style: NerfRenderer.render_style with the frame's latent, rendering.py:118-178).
    Yields (frame_index, {"rgb": [H*W,3], "depth": [H*W], "acc": [H*W]}) on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = H * W
    keys = ("rgb", "depth", "acc")

    def frame(i, b, cnt):
        if latents is None:
            return renderer.render_frame(H, W, K, poses[i], pix_begin=b, n=cnt, **render_kw)
        lat = latents if (torch.is_tensor(latents) and latents.dim() == 1) else latents[i]
        ro, rd = renderer.raygen(H, W, K, poses[i], pix_begin=b, n=cnt)
        return renderer.render_style(ro, rd, lat, **render_kw)

    if split == "rows" or world == 1:
        b, e = shard_range(n, rank, world)
        for i in range(len(poses)):
            out = frame(i, b, e - b)
            yield i, gather_tiles({k: out[k] for k in keys}, n, group)
        return
    if split != "frames":
        raise ValueError("split must be 'rows' or 'frames'")
    # whole frames round-robin: ONE packed all-gather per group of `world` frames, issued on a side stream so that the next
    # group's rendering overlaps it; a group's frames are yielded once the following group has been enqueued
    tg = TileGatherer(n, renderer.device, group)
    pending = None
    for base in list(range(0, len(poses), world)) + [None]:
        handle = None
        if base is not None:
            idx = min(base + rank, len(poses) - 1)          # ranks past the end re-render the last frame (dropped below)
            out = tg.outputs()
            if latents is None:
                renderer.render_frame(H, W, K, poses[idx], pix_begin=0, n=n, out=out, **render_kw)
            else:
                lat = latents if (torch.is_tensor(latents) and latents.dim() == 1) else latents[idx]
                ro, rd = renderer.raygen(H, W, K, poses[idx], pix_begin=0, n=n)
                renderer.render_style(ro, rd, lat, out=out, **render_kw)
            handle = (base, tg.gather_async())
        if pending is not None:
            pbase, h = pending
            full = h.wait()
            for r in range(world):
                if pbase + r < len(poses):
                    yield pbase + r, {k: full[k][r].clone() for k in keys}
        pending = handle
    tg.finish()
