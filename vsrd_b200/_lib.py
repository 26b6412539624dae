"""ctypes binding of libvsrd_b200.so (include/vsrd_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  The library is built in-tree by `python -m vsrd_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvsrd_b200.so")

MLP_WEIGHTS = 1617
GRAD_STRIDE = 1632
MAX_INSTANCES = 32
CULL_COUNT_STRIDE = 32                 # VSRD_CULL_COUNT_STRIDE
CULL_HEADER_INTS = MAX_INSTANCES * CULL_COUNT_STRIDE
MAX_INTERVALS = 512

c_float_p = ctypes.c_void_p  # device pointers travel as opaque addresses


class VsrdScene(ctypes.Structure):
    _fields_ = [
        ("num_instances", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("locations", ctypes.c_void_p),
        ("rotations", ctypes.c_void_p),
        ("half_extents", ctypes.c_void_p),
        ("mlp_weights", ctypes.c_void_p),
        ("temperature", ctypes.c_float),
        ("scale", ctypes.c_float),
        ("step_state", ctypes.c_void_p),
    ]


class VsrdStepState(ctypes.Structure):
    """Device-resident per-step scalars (this mirror is used for sizes/offsets and host read-back)."""
    _fields_ = [
        ("temperature", ctypes.c_float),
        ("std_deviation", ctypes.c_float),
        ("cosine_ratio", ctypes.c_float),
        ("eikonal_weight", ctypes.c_float),
        ("seed", ctypes.c_uint64),
        ("step", ctypes.c_int64),
    ]


class VsrdSchedule(ctypes.Structure):
    _fields_ = [
        ("num_steps", ctypes.c_int64),
        ("warmup_steps", ctypes.c_int64),
        ("max_temperature", ctypes.c_float),
        ("min_temperature", ctypes.c_float),
        ("max_std_deviation", ctypes.c_float),
        ("min_std_deviation", ctypes.c_float),
        ("eikonal_weight", ctypes.c_float),
        ("_pad", ctypes.c_float),
        ("seed", ctypes.c_uint64),
    ]


class VsrdViews(ctypes.Structure):
    _fields_ = [
        ("num_views", ctypes.c_int32),
        ("target_view", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("extrinsics", ctypes.c_void_p),
        ("intrinsics", ctypes.c_void_p),
    ]


class VsrdRays(ctypes.Structure):
    _fields_ = [
        ("num_rays", ctypes.c_int32),
        ("num_intervals", ctypes.c_int32),
        ("origins", ctypes.c_void_p),
        ("directions", ctypes.c_void_p),
        ("distances", ctypes.c_void_p),
        ("forward_samples", ctypes.c_void_p),
        ("cull_stats", ctypes.c_void_p),
        ("live_tiles", ctypes.c_void_p),
    ]


class VsrdRenderParams(ctypes.Structure):
    _fields_ = [
        ("std_deviation", ctypes.c_float),
        ("cosine_ratio", ctypes.c_float),
        ("epsilon", ctypes.c_float),
        ("_pad", ctypes.c_float),
    ]


class VsrdLoss(ctypes.Structure):
    _fields_ = [
        ("targets", ctypes.c_void_p),
        ("silhouette_weight", ctypes.c_float),
        ("eikonal_weight", ctypes.c_float),
    ]


HYPER_WIDTH = 256
HYPER_MAX_LAYERS = 5
MAX_PARAM_GROUPS = 8


class VsrdHyperLayer(ctypes.Structure):
    _fields_ = [
        ("weight_v", ctypes.c_void_p),
        ("weight_g", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("ln_weight", ctypes.c_void_p),
        ("ln_bias", ctypes.c_void_p),
        ("in_features", ctypes.c_int32),
        ("out_features", ctypes.c_int32),
    ]


class VsrdHyperNet(ctypes.Structure):
    _fields_ = [("num_layers", ctypes.c_int32), ("_pad", ctypes.c_int32), ("layers", VsrdHyperLayer * HYPER_MAX_LAYERS)]


class VsrdHyperLayerGrads(ctypes.Structure):
    _fields_ = [
        ("weight_v", ctypes.c_void_p),
        ("weight_g", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("ln_weight", ctypes.c_void_p),
        ("ln_bias", ctypes.c_void_p),
    ]


class VsrdHyperNetGrads(ctypes.Structure):
    _fields_ = [("num_layers", ctypes.c_int32), ("_pad", ctypes.c_int32), ("layers", VsrdHyperLayerGrads * HYPER_MAX_LAYERS)]


class VsrdBoxRanges(ctypes.Structure):
    _fields_ = [
        ("location_min", ctypes.c_float * 3),
        ("location_max", ctypes.c_float * 3),
        ("dimension_min", ctypes.c_float * 3),
        ("dimension_max", ctypes.c_float * 3),
    ]


class VsrdAdamGroups(ctypes.Structure):
    _fields_ = [
        ("num_groups", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("group_end", ctypes.c_int64 * MAX_PARAM_GROUPS),
        ("first_step", ctypes.c_int64 * MAX_PARAM_GROUPS),
        ("base_lr", ctypes.c_float * MAX_PARAM_GROUPS),
        ("beta1", ctypes.c_float),
        ("beta2", ctypes.c_float),
        ("eps", ctypes.c_float),
        ("_pad2", ctypes.c_float),
        ("log_gamma", ctypes.c_double),
    ]


# name -> (restype, argtypes); must list every symbol include/vsrd_b200.h declares
_P = ctypes.POINTER
_V = ctypes.c_void_p
_I = ctypes.c_int
SIGNATURES = {
    "vsrd_version": (_I, []),
    "vsrd_last_error": (ctypes.c_char_p, []),
    "vsrd_backward_blocks_per_instance": (_I, [_I, _I, _I]),
    "vsrd_backward_tile_rows": (_I, []),
    "vsrd_ray_directions": (_I, [_V, _I, _I, _I, _V, _V]),
    "vsrd_gather_rays": (_I, [_V, _V, _V, _I, _I, _I, _I, _V, _V, _V]),
    "vsrd_place_coarse": (_I, [_V, _V, ctypes.c_uint64, _V, _I, _I, _V, _V]),
    "vsrd_place_fine": (_I, [_V, _V, _V, ctypes.c_uint64, _V, _I, _I, _V, _V]),
    "vsrd_cull_samples": (_I, [_P(VsrdScene), _P(VsrdRays), _V, _V, _V]),
    "vsrd_live_tiles_bytes": (ctypes.c_size_t, [_I, _I, _I]),
    "vsrd_field_forward": (_I, [_P(VsrdScene), _P(VsrdRays), _V, _V]),
    "vsrd_composite_forward": (_I, [_P(VsrdScene), _P(VsrdRays), _P(VsrdRenderParams), _V, _V, _V, _V,
                                    _P(VsrdLoss), _V, _V]),
    "vsrd_composite_backward": (_I, [_P(VsrdScene), _P(VsrdRays), _P(VsrdRenderParams), _V, _V, _V, _V,
                                     _P(VsrdLoss), _V, _V, _V]),
    "vsrd_field_backward": (_I, [_P(VsrdScene), _P(VsrdRays), _V, _V, _V, _V, _V, _V, _V]),
    "vsrd_experimental_field_backward_tcgen05": (_I, [_P(VsrdScene), _P(VsrdRays), _V, _V, _V, _V, _V, _V, _V]),
    "vsrd_projection_scratch_floats": (ctypes.c_size_t, [_I, _I]),
    "vsrd_project_box_3d": (_I, [_V, _I, _V, ctypes.c_float, _V, _V]),
    "vsrd_project_box_3d_backward": (_I, [_V, _I, _V, ctypes.c_float, _V, _V, _V]),
    "vsrd_projection_step": (_I, [_P(VsrdViews), _I, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    "vsrd_ray_cdf_scratch_doubles": (ctypes.c_size_t, [ctypes.c_int64]),
    "vsrd_ray_cdf_build": (_I, [_V, ctypes.c_int64, _I, _V, _V, _V]),
    "vsrd_select_rays": (_I, [_V, ctypes.c_int64, _V, _I, ctypes.c_uint64, _V, _I, _V, _V, _V]),
    "vsrd_gather_targets": (_I, [_V, _V, _V, _I, _I, _V, _V]),
    "vsrd_soft_masks": (_I, [_V, _V, _I, _I, _I, _I, _I, ctypes.c_float, _V, _V]),
    "vsrd_field_points": (_I, [_P(VsrdScene), _V, _I, _V, _V]),
    "vsrd_union_points": (_I, [_P(VsrdScene), _V, _I, _V, _V, _V]),
    "vsrd_sphere_trace_step": (_I, [_V, _V, _I, _I, ctypes.c_float, ctypes.c_float, _V, _V, _V, _V, _I, _V]),
    "vsrd_hyper_scratch_floats": (ctypes.c_size_t, [_I]),
    "vsrd_hyper_forward": (_I, [_P(VsrdHyperNet), _V, _I, _V, _V, _V]),
    "vsrd_hyper_backward": (_I, [_P(VsrdHyperNet), _P(VsrdHyperNetGrads), _V, _I, _V, _V, _V, _V, _V]),
    "vsrd_decode_boxes": (_I, [_P(VsrdBoxRanges), _V, _V, _V, _I, _V, _V, _V, _V, _V]),
    "vsrd_decode_boxes_backward": (_I, [_P(VsrdBoxRanges), _V, _V, _V, _I, _V, _V, _V, _V, _V, _V, ctypes.c_float,
                                        ctypes.c_float, _V, _V, _V, _V, _V, _V, _V]),
    "vsrd_adam_step": (_I, [_V, _V, _V, _V, ctypes.c_int64, _P(VsrdAdamGroups), _V, ctypes.c_int64, _V]),
    "vsrd_step_state_update": (_I, [_V, _P(VsrdSchedule), ctypes.c_int64, _V]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"vsrd_b200: {LIB_PATH} not found. Build it with `python -m vsrd_b200.build` "
            "(needs nvcc, targets sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().vsrd_last_error()
        raise RuntimeError(msg.decode() if msg else f"vsrd_b200: call failed with status {status}")
