"""The 120-frame stylised spiral path (BASELINE config 3: render_valid_style at 1008x756) on the GPUs of one box, whole frames
round-robin over ranks, one NCCL tile all-gather per group of `world` frames.  Prints one JSON line (rank 0).
    python tools/render_spiral.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/render_spiral.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench as B
import tgtc_style_b200 as T
from bench import spiral_poses, H, W, FOCAL


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wc, wf = B.synth_nerf_weights(0)
    cs, ws = B.synth_style_weights(1)
    r = T.NerfRenderer(device=dev, mode="bf16")
    r.set_weights(wc, wf)
    r.set_style_weights(cs, ws)
    K = np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]])
    poses = spiral_poses(120)[:nframes]
    table = torch.randn(20, 32, generator=torch.Generator().manual_seed(3)).repeat(7, 1)[:nframes].to(dev)   # models.py:496 tiles 20 latents x7
    for _ in T.render_path_sharded(r, H, W, K, poses[:world], split="frames", latents=table[:world], chunk=32768):
        pass                                    # warm-up group
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    checksum, got = 0.0, 0
    for i, fr in T.render_path_sharded(r, H, W, K, poses, split="frames", latents=table, chunk=32768):
        got += 1
        if i % 40 == 0:
            checksum += float(fr["rgb"].double().sum())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.time() - t0
    if rank == 0:
        print(json.dumps({"workload": "120-frame stylised spiral (render_valid_style) at 1008x756, frames round-robin over ranks, "
                                      "NCCL tile all-gather per frame group", "n_gpus": world, "frames": got, "seconds": dt,
                          "frames_per_s": got / dt, "rays_per_s": got * H * W / dt, "checksum": checksum}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
