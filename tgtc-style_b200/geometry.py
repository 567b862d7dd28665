"""cal_geometry's outputs (rendering.py:5-90), emitted from device buffers.

The reference renders every training pose with the plain NeRF chain, derives per pixel the NDC surface point
coor = t_exp * rays_d + rays_o (rendering.py:54) and writes, per frame, rgb_%05d.png, a min-max normalised
depth_%05d.png and geometry_%05d.npz {coor_map [H,W,3], cps, hwf, near, far} -- the files the 2-D temporal trainer
(train_style_modules.py:101-115, :445-451) consumes.  Here the frame is rendered on the device from the pose
(NerfRenderer.render_frame), coor_map is formed there, and one D2H copy per frame feeds the writers.
"""
import os

import numpy as np
import torch


def frame_geometry(renderer, H, W, K, pose, near=0., far=1., **render_kw):
    """-> {"rgb" [H,W,3], "t" [H,W], "coor_map" [H,W,3]} device tensors for one pose (rendering.py:27-54)."""
    ro, rd = renderer.raygen(H, W, K, pose)
    out = renderer.render(ro, rd, near, far, want_weights=False, **render_kw)
    coor = out["depth"].unsqueeze(-1) * rd + ro                      # rendering.py:54
    return {"rgb": out["rgb"].view(H, W, 3), "t": out["depth"].view(H, W), "coor_map": coor.view(H, W, 3)}


def _to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)                  # utils.to8b


def _imwrite(path, img):
    try:
        import cv2
        cv2.imwrite(path, img[..., ::-1] if img.ndim == 3 else img)
    except ImportError:                                               # pragma: no cover
        from PIL import Image
        Image.fromarray(img).save(path)


def save_frame(sv_path, index, geom, cps_i, hwf, near, far):
    """rendering.py:66-76 for one frame: rgb_%05d.png, depth_%05d.png (per-frame min-max), geometry_%05d.npz."""
    os.makedirs(sv_path, exist_ok=True)
    rgb = geom["rgb"].detach().cpu().numpy().astype(np.float32)
    t = geom["t"].detach().cpu().numpy().astype(np.float32)
    coor = geom["coor_map"].detach().cpu().numpy().astype(np.float32)
    tn = (t - t.min()) / (t.max() - t.min() + 1e-7)
    # the reference converts to int32 in [0,255] and then clips through to8b (rendering.py:72-75): white-saturated PNGs;
    # the images are written here from the [0,1] values (the geometry npz, which the consumer reads, is identical)
    _imwrite(os.path.join(sv_path, "rgb_%05d.png" % index), _to8b(rgb))
    _imwrite(os.path.join(sv_path, "depth_%05d.png" % index), _to8b(tn))
    np.savez(os.path.join(sv_path, "geometry_%05d" % index), coor_map=coor, cps=np.asarray(cps_i), hwf=np.asarray(hwf), near=near, far=far)
    return coor


def cal_geometry(renderer, H, W, K, poses, hwf, near=0., far=1., sv_path=None, **render_kw):
    """The frame loop of rendering.cal_geometry over `poses` (c2w 3x4 each): returns (rgb_map [F,H,W,3], t_map [F,H,W,1])
    as numpy like the reference and, with sv_path, writes the per-frame files plus geometry.npz (rendering.py:78-81)."""
    rgbs, ts, coors = [], [], []
    for i, pose in enumerate(poses):
        g = frame_geometry(renderer, H, W, K, pose, near, far, **render_kw)
        if sv_path is not None:
            coors.append(save_frame(sv_path, i, g, pose, hwf, near, far))
        else:
            coors.append(g["coor_map"].cpu().numpy())
        rgbs.append(g["rgb"].cpu().numpy())
        ts.append(g["t"].cpu().numpy()[..., None])
    if sv_path is not None:
        np.savez(os.path.join(sv_path, "geometry"), coor_map=np.stack(coors), cps=np.asarray(poses), hwf=np.asarray(hwf), near=near, far=far)
    return np.stack(rgbs), np.stack(ts)
