"""Shared pieces of the drop-in tests: a tiny stand-in for the reference's dataset / DataLoader (the reference's own classes
need LLFF images on disk) and the arguments cal_geometry reads."""
import numpy as np
import torch

import render_oracle as O


class Args:
    N_samples = 64
    N_samples_fine = 64
    chunk = 1024
    sigma_noise_std = 1.0


class FakeDataset:
    """the attributes rendering.cal_geometry reads (rendering.py:8-12): cps / cps_valid, hwf, near, far, frame_num, h, w, mode"""

    def __init__(self, H, W, f, poses):
        self.h, self.w, self.hwf = H, W, [H, W, f]
        self.cps = np.stack([np.concatenate([p, [[0, 0, 0, 1]]], 0) for p in poses], 0)
        self.cps_valid = self.cps
        self.frame_num = len(poses)
        self.near, self.far = 0., 1.
        self.mode = "train"
        rays = [O.make_rays(H, W, f, p) for p in poses]
        self.rays_o = np.concatenate([r[0] for r in rays], 0)
        self.rays_d = np.concatenate([r[1] for r in rays], 0)


class FakeLoader:
    """iterates batches the way the reference's DataLoader does: dicts of CPU tensors (rendering.py:18-21 calls .numpy())"""

    def __init__(self, dataset, batch_size):
        self.dataset, self.bs = dataset, batch_size

    def __len__(self):
        return (self.dataset.rays_o.shape[0] + self.bs - 1) // self.bs

    def __iter__(self):
        d = self.dataset
        for i in range(0, d.rays_o.shape[0], self.bs):
            yield {"rays_o": torch.from_numpy(d.rays_o[i:i + self.bs]), "rays_d": torch.from_numpy(d.rays_d[i:i + self.bs])}


POSES = [np.eye(4)[:3, :4], np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.05], [0, 0, 1, 0.02]], dtype=np.float64)]
