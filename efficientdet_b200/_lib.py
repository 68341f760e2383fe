"""ctypes binding of libeffdet_b200.so (the C ABI declared in include/effdet_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# EFFDET_B200_LIB: another build of the same library (A/B timing of two kernel versions inside one GPU job)
LIB_PATH = os.environ.get("EFFDET_B200_LIB") or os.path.join(HERE, "libeffdet_b200.so")

OK, E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED = 0, -1, -2, -3, -4
F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SWISH, ACT_SIGMOID = 0, 1, 2, 3


class EffdetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("effdet_b200 error %d: %s" % (code, msg))
        self.code = code


class CapacityError(EffdetError):
    pass


_lib = None

c_void_p, c_int, c_size_t, c_float, c_double = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t,
                                                ctypes.c_float, ctypes.c_double)
_F4 = c_float * 4

_SIGNATURES = {
    "effdet_anchors_for_shape_host": [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                      c_void_p, c_int, c_void_p, c_size_t],
    "effdet_compute_overlap": [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p],
    "effdet_anchor_targets": [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_void_p, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p],
    "effdet_regress_boxes": [c_void_p, c_int, c_void_p, _F4, _F4, c_int, c_size_t, c_void_p,
                             c_void_p],
    "effdet_clip_boxes": [c_void_p, c_int, c_size_t, c_float, c_float, c_void_p, c_void_p],
    "effdet_regress_clip_boxes": [c_void_p, c_int, c_void_p, _F4, _F4, c_int, c_size_t, c_float,
                                  c_float, c_void_p, c_void_p],
    "effdet_zero": [c_void_p, c_size_t, c_void_p],
    "effdet_bn_fold": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                       c_void_p],
    "effdet_stem_conv": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                         c_int, c_int, c_void_p],
    "effdet_stem_conv_act": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                             c_int, c_int, c_void_p],
    "effdet_stem_conv_u8": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_void_p],
    "effdet_replay_load": [ctypes.c_char_p, c_int, c_void_p],
    "effdet_replay_region": [c_void_p, ctypes.c_char_p, c_void_p, c_void_p],
    "effdet_replay_step": [c_void_p, c_double, c_void_p],
    "effdet_replay_destroy": [c_void_p],
    "effdet_normalize_u8": [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "effdet_letterbox_geometry": [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "effdet_letterbox_u8": [c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_int, c_void_p],
    "effdet_conv2d": [c_void_p, c_void_p],
    "effdet_conv_weight_panel": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p],
    "effdet_split_bf16": [c_void_p, c_void_p, c_size_t, c_int, c_void_p],
    "effdet_dwconv_split_out": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_conv_weight_panel_split": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "effdet_dwconv": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_se_gate": [c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                       c_int, c_int, c_int, c_void_p],
    "effdet_wbifpn_add": [c_void_p, c_int, c_void_p, c_float, c_void_p, c_size_t, c_int, c_void_p],
    "effdet_bifpn_node": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                          c_void_p],
    "effdet_detection_losses": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_size_t, c_int, c_float, c_float, c_float, c_float, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_void_p],
    "effdet_resample_fuse": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_bn_train_stats": [c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_float, c_float,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_void_p],
    "effdet_dwconv_bn_stats": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "effdet_scale_shift_act": [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int,
                               c_int, c_void_p],
    "effdet_bn_relu_backward": [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_void_p],
    "effdet_colsum": [c_void_p, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                      c_void_p],
    "effdet_colsum_tail": [c_void_p, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p],
    "effdet_dw_wgrad": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                        c_int, c_void_p],
    "effdet_fuse_backward_input": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_float,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_fuse_backward_weights": [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_int, c_void_p],
    "effdet_conv_wgrad": [c_void_p, c_void_p],
    "effdet_conv_wgrad_tc": [c_void_p, c_void_p],
    "effdet_conv_dgrad_strided": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_void_p],
    "effdet_conv_weight_transpose": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "effdet_flip_taps": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "effdet_zero_insert": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_void_p],
    "effdet_sgd_momentum_step": [c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float,
                                 c_void_p],
    "effdet_sgd_momentum_step_dev_lr": [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_float, c_float, c_void_p],
    "effdet_bn_act_backward": [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_void_p],
    "effdet_spatial_sum": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_se_apply": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_drop_connect_scales": [c_void_p, c_int, c_int, ctypes.c_ulonglong, c_void_p, c_void_p, c_void_p],
    "effdet_sample_scale_add": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_size_t, c_int, c_void_p],
    "effdet_se_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                           c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_se_bn_backward": [c_void_p] * 4 + [c_int] + [c_void_p] * 8 + [c_void_p] * 5 + [c_void_p] * 2 +
                             [c_void_p] * 2 + [c_void_p] * 3 + [c_int] + [c_void_p] * 2 + [c_int] * 5 + [c_void_p],
    "effdet_dw_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_int, c_int, c_int, c_int, c_int, c_void_p],
    "effdet_stem_wgrad": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_void_p],
    "effdet_filter_detections": [c_void_p, c_void_p, c_int, c_size_t, c_int, c_float, c_float,
                                 c_int, c_int, c_int, c_void_p, c_size_t, c_size_t, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    # plan level (csrc/plan.cu)
    "effdet_plan_create": [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_uint, c_void_p],
    "effdet_plan_destroy": [c_void_p],
    "effdet_plan_weight_info": [c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "effdet_plan_bind_weights": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "effdet_plan_bind_weights_host": [c_void_p, c_void_p, c_void_p, c_int],
    "effdet_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "effdet_plan_buffers": [c_void_p, c_void_p, c_void_p, c_void_p],
    "effdet_detect": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                      c_void_p],
    "effdet_detect_host": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p],
}


MAX_GROUPS = 5


class ConvDesc(ctypes.Structure):
    """Mirror of `effdet_conv_desc` (include/effdet_b200.h)."""
    _fields_ = [
        ("n_groups", c_int),
        ("x", c_void_p * MAX_GROUPS), ("y", c_void_p * MAX_GROUPS),
        ("residual", c_void_p * MAX_GROUPS),
        ("H", c_int * MAX_GROUPS), ("W", c_int * MAX_GROUPS), ("ldc", c_int * MAX_GROUPS),
        ("y_batch_stride", ctypes.c_longlong * MAX_GROUPS),
        ("ldx", c_int * MAX_GROUPS), ("x_batch_stride", ctypes.c_longlong * MAX_GROUPS),
        ("relu_mask", c_void_p * MAX_GROUPS),
        ("B", c_int), ("Cin", c_int), ("Cout", c_int), ("kh", c_int), ("kw", c_int),
        ("stride", c_int),
        ("weight", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("gate", c_void_p),
        ("keep", c_void_p),
        ("act", c_int), ("in_dtype", c_int), ("out_dtype", c_int),
        ("weight_bf16", c_void_p), ("allow_tensor_core", c_int), ("weight_per_sample", c_int),
        ("split_planes", c_int),
    ]


class WgradDesc(ctypes.Structure):
    """Mirror of `effdet_wgrad_desc` (include/effdet_b200.h)."""
    _fields_ = [
        ("n_groups", c_int),
        ("x", c_void_p * MAX_GROUPS), ("dz", c_void_p * MAX_GROUPS),
        ("H", c_int * MAX_GROUPS), ("W", c_int * MAX_GROUPS), ("dz_ld", c_int * MAX_GROUPS),
        ("dz_batch_stride", ctypes.c_longlong * MAX_GROUPS),
        ("B", c_int), ("Cin", c_int), ("Cout", c_int), ("kh", c_int), ("kw", c_int),
        ("stride", c_int),
        ("dweight", c_void_p), ("partial", c_void_p), ("n_splits", c_int), ("accumulate", c_int),
        ("x_dtype", c_int), ("dz_dtype", c_int), ("dbias", c_void_p),
    ]


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "efficientdet_b200: %s is missing. Build it with `python -m efficientdet_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.effdet_last_error.restype = ctypes.c_char_p
    lib.effdet_version.restype = c_int
    lib.effdet_launch_count.restype = ctypes.c_longlong
    lib.effdet_filter_detections_workspace_size.restype = c_size_t
    lib.effdet_filter_detections_workspace_size.argtypes = [c_int, c_size_t, c_int, c_size_t, c_int]
    lib.effdet_detection_losses_workspace_size.restype = c_size_t
    lib.effdet_detection_losses_workspace_size.argtypes = []
    lib.effdet_colreduce_blocks.restype = c_int
    lib.effdet_colreduce_blocks.argtypes = [c_size_t, c_int, c_int]
    lib.effdet_dw_wgrad_blocks.restype = c_int
    lib.effdet_dw_wgrad_blocks.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.effdet_conv_wgrad_splits.restype = c_int
    lib.effdet_conv_wgrad_splits.argtypes = [c_void_p]
    lib.effdet_conv_wgrad_tc_splits.restype = c_int
    lib.effdet_conv_wgrad_tc_splits.argtypes = [c_void_p]
    lib.effdet_conv_tc_block_n.restype = c_int
    lib.effdet_conv_tc_block_n.argtypes = [c_int]
    lib.effdet_conv_weight_panel_elems.restype = c_size_t
    lib.effdet_conv_weight_panel_elems.argtypes = [c_int, c_int, c_int]
    lib.effdet_conv_weight_panel_split_elems.restype = c_size_t
    lib.effdet_conv_weight_panel_split_elems.argtypes = [c_int, c_int, c_int]
    lib.effdet_se_backward_blocks.restype = c_int
    lib.effdet_se_backward_blocks.argtypes = [c_int, c_int, c_int]
    lib.effdet_se_bn_backward_blocks.restype = c_int
    lib.effdet_se_bn_backward_blocks.argtypes = [c_int, c_int, c_int, c_int]
    lib.effdet_dw_backward_blocks.restype = c_int
    lib.effdet_dw_backward_blocks.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, c_int]
    lib.effdet_stem_wgrad_blocks.restype = c_int
    lib.effdet_stem_wgrad_blocks.argtypes = [c_int, c_int, c_int]
    lib.effdet_dwconv_se_blocks.restype = c_int
    lib.effdet_dwconv_se_blocks.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
    lib.effdet_plan_num_anchors.restype = c_size_t
    lib.effdet_plan_num_anchors.argtypes = [c_void_p]
    lib.effdet_plan_num_weights.restype = c_int
    lib.effdet_plan_num_weights.argtypes = [c_void_p]
    lib.effdet_plan_num_launches.restype = c_int
    lib.effdet_plan_num_launches.argtypes = [c_void_p]
    lib.effdet_plan_dry_run.restype = c_int
    lib.effdet_plan_dry_run.argtypes = [c_void_p, c_void_p]
    lib.effdet_replay_num_launches.restype = c_int
    lib.effdet_replay_num_launches.argtypes = [c_void_p]
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    _lib = lib
    return lib


def register(name, argtypes, restype=c_int):
    """Used by sibling modules to declare further entry points."""
    _SIGNATURES[name] = argtypes
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = restype


def check(code):
    if code == OK:
        return
    msg = load().effdet_last_error().decode("utf-8", "replace")
    if code == E_CAPACITY:
        raise CapacityError(code, msg)
    if code == E_INVALID:
        raise ValueError("effdet_b200: " + msg)
    raise EffdetError(code, msg)


def call(name, *args):
    check(getattr(load(), name)(*args))


def launch_count():
    return int(load().effdet_launch_count())


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
