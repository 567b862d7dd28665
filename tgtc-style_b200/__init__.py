"""tgtc-style_b200 -- B200-native (sm_100a) NeRF ray-render hot path of TGTC-Style.

Public surface:
  NerfRenderer            render(rays_o, rays_d, near, far, chunk) -> {rgb, depth, acc, weights}
  make_callables / patch  the reference's four injected callables, B200-backed
  NerfTrainer             data-parallel training step (forward + backward + Adam), one gradient all-reduce
  StyleTrainer            Style_train iteration: style modules + per-ray latents on frozen NeRF nets
  shard_range / gather_tiles / TileGatherer / render_frame_sharded   multi-GPU ray sharding; the tile all-gather off the critical path
The directory name contains a hyphen; import it as `tgtc_style_b200` (root-level loader module).
"""
from . import _lib
from ._lib import MLP_BF16, MLP_F16, MLP_FP32, NET_COARSE, NET_FINE, TgtcError
from .dist import TileGatherer, gather_tiles, render_frame_sharded, render_path_sharded, shard_range, shard_sizes
from .geometry import cal_geometry, frame_geometry, save_frame
from .render import LAYER_NAMES, LAYER_SHAPES, NerfRenderer
from .shims import make_callables, patch
from .train import NerfTrainer, StyleLatents, StyleTrainer

__all__ = ["NerfRenderer", "NerfTrainer", "StyleTrainer", "StyleLatents", "make_callables", "patch", "shard_range", "shard_sizes", "gather_tiles", "TileGatherer", "render_frame_sharded", "render_path_sharded",
           "cal_geometry", "frame_geometry", "save_frame", "TgtcError", "MLP_FP32", "MLP_BF16", "MLP_F16", "NET_COARSE", "NET_FINE", "LAYER_NAMES", "LAYER_SHAPES"]
