"""ctypes binding of libtgtc_b200.so (the C ABI of include/tgtc_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this
raises.  PyTorch is used by the callers only to own device memory and streams.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TGTC_B200_LIB: another build of the same library (development A/B runs on one box: tools/ab_builds.py)
LIB_PATH = os.environ.get("TGTC_B200_LIB") or os.path.join(HERE, "libtgtc_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64

MLP_FP32 = 0
MLP_BF16 = 1
MLP_F16 = 2
NET_COARSE = 0
NET_FINE = 1
NUM_PARAMS = 24

OUT_FIELDS = ("rgb", "depth", "acc", "weights", "rgb_coarse", "depth_coarse", "acc_coarse", "weights_coarse", "ts_fine")


class RenderOut(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in OUT_FIELDS]


# name -> (restype, argtypes); every symbol include/tgtc_b200.h declares
PROTOTYPES = {
    "tgtc_last_error": (ctypes.c_char_p, []),
    "tgtc_abi_version": (ctypes.c_int, []),
    "tgtc_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "tgtc_destroy": (ctypes.c_int, [c_void_p]),
    "tgtc_set_weights": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.POINTER(c_void_p), c_void_p]),
    "tgtc_raygen": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_int, c_i64, c_i64, c_void_p, c_void_p, c_void_p]),
    "tgtc_sample_uniform": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgtc_nerf_forward": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, ctypes.c_int, c_i64, ctypes.c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgtc_nerf_forward_rays": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_double, c_void_p, c_void_p]),
    "tgtc_composite": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, ctypes.c_int, c_i64,
                                      ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgtc_composite_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, ctypes.c_int, c_i64, ctypes.c_int,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgtc_sample_fine": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, ctypes.c_int, ctypes.c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tgtc_render_workspace_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render_workspace_bytes_mode": (ctypes.c_size_t, [ctypes.c_int, c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render_frame_workspace_bytes_mode": (ctypes.c_size_t, [ctypes.c_int, c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                   ctypes.c_int, c_i64, ctypes.c_int, ctypes.POINTER(RenderOut), c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_render_host": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_int, c_i64, ctypes.c_int, ctypes.POINTER(RenderOut), c_void_p]),
    "tgtc_render_frame_workspace_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render_frame": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int,
                                         ctypes.c_double, ctypes.c_int, c_i64, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                         c_i64, ctypes.c_int, ctypes.POINTER(RenderOut), c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_train_set_coarse_event": (ctypes.c_int, [c_void_p, c_void_p]),
    "tgtc_train_workspace_bytes": (ctypes.c_size_t, [c_void_p, c_i64, ctypes.c_int, ctypes.c_int]),
    "tgtc_num_params": (c_i64, []),
    "tgtc_train_step": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_adam_step": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_double, c_i64, c_void_p]),
    "tgtc_set_style_weights": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_void_p), c_void_p]),
    "tgtc_render_style_workspace_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render_style": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                         ctypes.c_int, c_i64, c_void_p, c_void_p, ctypes.POINTER(RenderOut), c_void_p,
                                         ctypes.c_size_t, c_void_p]),
    "tgtc_train_step_seeded": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_double, c_void_p,
                                              ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_philox_fill": (ctypes.c_int, [c_void_p, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_i64, c_void_p, c_void_p]),
    "tgtc_style_train_workspace_bytes": (ctypes.c_size_t, [c_void_p, c_i64, ctypes.c_int, ctypes.c_int]),
    "tgtc_style_num_params": (c_i64, []),
    "tgtc_style_train_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                ctypes.c_size_t, c_void_p]),
    "tgtc_style_train_backward": (ctypes.c_int, [c_void_p, c_i64, ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p, ctypes.c_size_t,
                                                 c_void_p]),
    "tgtc_style_train_forward_seeded": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                       ctypes.c_int, c_void_p, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_double, c_void_p,
                                                       c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_style_train_backward_seeded": (ctypes.c_int, [c_void_p, c_i64, ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_ulonglong,
                                                        ctypes.c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p, ctypes.c_int,
                                                        c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_style_loss_sums": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_i64, c_void_p, c_void_p]),
    "tgtc_style_loss_grads": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_i64, c_void_p, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "tgtc_style_latents_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "tgtc_style_latents_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_double, c_void_p, ctypes.c_double, c_void_p,
                                                   ctypes.c_int, c_void_p]),
    "tgtc_nerf_stash_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int]),
    "tgtc_nerf_backward_scratch_bytes": (ctypes.c_size_t, [c_void_p, c_i64, ctypes.c_int]),
    "tgtc_nerf_forward_stash": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, ctypes.c_int, c_void_p, c_void_p,
                                               ctypes.c_size_t, c_void_p]),
    "tgtc_nerf_backward": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_i64, ctypes.c_int, c_void_p, c_void_p, c_void_p, ctypes.c_int,
                                          c_void_p, ctypes.c_size_t, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_render_style_rays_workspace_bytes": (ctypes.c_size_t, [c_i64, ctypes.c_int, ctypes.c_int, c_i64]),
    "tgtc_render_style_rays": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                              ctypes.c_int, c_i64, c_void_p, ctypes.POINTER(RenderOut), c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_style_stage_workspace_bytes": (ctypes.c_size_t, [c_i64]),
    "tgtc_style_concat_forward": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "tgtc_style_forward": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, ctypes.c_size_t,
                                          c_void_p]),
    "tgtc_profile_enable": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "tgtc_profile_read": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_i64), c_double_p, c_double_p]),
    "tgtc_profile_read_kind": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.POINTER(c_i64), c_double_p, c_double_p]),
    "tgtc_launch_count": (c_i64, [c_void_p]),
}

# test hook exported by mlp_tc.cu (not in the public header)
DEBUG_PROTOTYPES = {
    "tgtc_debug_tc_layers": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_i64, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_double, ctypes.c_int, c_void_p, c_void_p]),
    "tgtc_debug_tc_f16": (None, [ctypes.c_int]),
    "tgtc_debug_no_fused_composite": (None, [ctypes.c_int]),
    "tgtc_debug_no_fused_sample_fine": (None, [ctypes.c_int]),
    "tgtc_debug_sample_fine_general": (None, [ctypes.c_int]),
}

_lib = None


class TgtcError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TgtcError(
            "libtgtc_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in DEBUG_PROTOTYPES.items():      # test hooks: an older build under TGTC_B200_LIB may lack the newest
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    if lib.tgtc_abi_version() != 2:
        raise TgtcError("libtgtc_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().tgtc_last_error()
        raise TgtcError("tgtc status %d: %s" % (status, msg.decode("utf-8", "replace") if msg else "?"))
