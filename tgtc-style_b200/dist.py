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
    for base in range(0, len(poses), world):
        mine = base + rank
        idx = min(mine, len(poses) - 1)          # ranks past the end re-render the last frame (dropped below)
        out = frame(idx, 0, n)
        full = gather_tiles({k: out[k] for k in keys}, n * world, group)   # equal shards: one all_gather_into_tensor per key
        for r in range(world):
            if base + r < len(poses):
                yield base + r, {k: full[k][r * n:(r + 1) * n] for k in keys}
