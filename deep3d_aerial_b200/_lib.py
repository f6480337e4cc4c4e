"""ctypes binding of libd3dsweep.so (declarations mirror include/d3d_sweep.h one to one).

There is no fallback: if the library is missing `load()` raises, and every op in this package
goes through it.  Build it with `python -m deep3d_aerial_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# D3D_SWEEP_LIB: an A/B build of the same library (tools/build_ab.py); there is no other implementation to point it at
LIB_PATH = os.environ.get("D3D_SWEEP_LIB") or os.path.join(HERE, "libd3dsweep.so")

# enums (include/d3d_sweep.h)
OK, ERR_BAD_ARGUMENT, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
AGG_WARP, AGG_VARIANCE, AGG_GROUP_CORR, AGG_WEIGHTED_PRODUCT, AGG_PAIR_MEAN = 0, 1, 2, 3, 4
SOFTMAX_STABLE, SOFTMAX_RAW_EXP, SOFTMAX_NONE = 0, 1, 2
CONF_MAX_PROB, CONF_WINDOW4 = 0, 1
HYPS_UNIFORM, HYPS_PER_PIXEL, HYPS_RESIZED = 0, 1, 2
SAMPLES_RANGE, SAMPLES_AROUND, SAMPLES_CASCADE, SAMPLES_SPREAD = 0, 1, 2, 3

_f32p = C.c_void_p  # device pointers travel as integers


class CostVolumeArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("mode", C.c_int32), ("num_views", C.c_int32), ("channels", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32), ("num_depth", C.c_int32),
        ("d_begin", C.c_int32), ("d_count", C.c_int32), ("hyps_per_pixel", C.c_int32),
        ("groups", C.c_int32), ("eps_in_numerator", C.c_int32), ("variant", C.c_int32), ("texel_slots", C.c_int32),
        ("feats", _f32p), ("pose", _f32p), ("hyps", _f32p), ("weights", _f32p), ("out", _f32p),
        ("out_stride_c", C.c_int64), ("out_stride_d", C.c_int64),
        ("rays", _f32p),
        ("view_slot", C.c_int32 * 9), ("reserved1", C.c_int32),
    ]


class RegressArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_depth", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("d_begin", C.c_int32), ("d_count", C.c_int32), ("softmax_mode", C.c_int32), ("conf_mode", C.c_int32),
        ("hyps_mode", C.c_int32), ("hyps_height", C.c_int32), ("hyps_width", C.c_int32),
        ("finalize", C.c_int32), ("next_num_depth", C.c_int32), ("lamb", C.c_float), ("reserved0", C.c_int32),
        ("next_interval", C.c_double),
        ("logits", _f32p), ("logits_stride_d", C.c_int64), ("hyps", _f32p),
        ("depth", _f32p), ("conf", _f32p), ("index", _f32p), ("state", _f32p),
        ("exp_variance", _f32p), ("next_hyps", _f32p),
        ("logit_planes", C.c_void_p * 16),
    ]


class SamplesArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("mode", C.c_int32), ("num_depth", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32), ("src_height", C.c_int32), ("src_width", C.c_int32),
        ("full_height", C.c_int32), ("full_width", C.c_int32),
        ("dmin", C.c_float), ("dmax", C.c_float), ("reserved0", C.c_int32),
        ("interval", C.c_double),
        ("cur", _f32p), ("out", _f32p), ("spread", _f32p),
    ]


FUSE_MAX_SRC, FUSE_GEOM_DOUBLES = 16, 64
REGRESS_MAX_PLANES = 16


class FuseArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_src", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("src_height", C.c_int32), ("src_width", C.c_int32), ("min_consistent", C.c_int32), ("accumulate", C.c_int32),
        ("position_threshold", C.c_double),
        ("depth_threshold", C.c_float), ("confidence_threshold", C.c_float), ("normal_threshold_cos", C.c_float),
        ("reserved1", C.c_int32),
        ("depth_ref", _f32p), ("normal_ref", _f32p), ("prob_ref", _f32p), ("geometry", C.c_void_p),
        ("depth_src", C.c_void_p * FUSE_MAX_SRC), ("normal_src", C.c_void_p * FUSE_MAX_SRC),
        ("depth_src_out", C.c_void_p * FUSE_MAX_SRC),
        ("mask", C.c_void_p), ("depth_reprojected", _f32p), ("xyz_world_src", _f32p), ("angle_conf", _f32p),
        ("consistent_count", C.c_void_p), ("xyz_fused", _f32p), ("final_mask", C.c_void_p),
        ("depth_ref_filtered", _f32p), ("accum", _f32p),
    ]


EXPORTS = {
    "d3d_cost_volume": (C.c_int, [C.POINTER(CostVolumeArgs), C.c_void_p]),
    "d3d_depth_regress": (C.c_int, [C.POINTER(RegressArgs), C.c_void_p]),
    "d3d_depth_samples": (C.c_int, [C.POINTER(SamplesArgs), C.c_void_p]),
    "d3d_consistency_fuse": (C.c_int, [C.POINTER(FuseArgs), C.c_void_p]),
    "d3d_pixel_rays": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "d3d_resize_bilinear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "d3d_homo_warp_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "d3d_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "d3d_last_error": (C.c_char_p, []),
    "d3d_version": (C.c_int, []),
    "d3d_launch_count": (C.c_int64, []),
    "d3d_abi_sizeof": (C.c_int32, [C.c_int32]),
}

_lib = None


class SweepError(RuntimeError):
    """A libd3dsweep call returned a non-zero status."""

    def __init__(self, code, message):
        super().__init__("libd3dsweep error %d: %s" % (code, message))
        self.code = code


def load() -> C.CDLL:
    """dlopen libd3dsweep.so and type its exports; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "libd3dsweep.so is missing at %s -- run `python -m deep3d_aerial_b200.build`; "
                "this package has no CPU or PyTorch fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        for which, struct in enumerate((CostVolumeArgs, RegressArgs, SamplesArgs, FuseArgs)):
            if lib.d3d_abi_sizeof(which) != C.sizeof(struct):
                raise ImportError("ABI mismatch: %s is %d bytes here, %d in libd3dsweep.so"
                                  % (struct.__name__, C.sizeof(struct), lib.d3d_abi_sizeof(which)))
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise SweepError(status, load().d3d_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().d3d_launch_count())
