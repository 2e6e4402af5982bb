"""ctypes binding of libts2d.so -- the C ABI declared in include/ts2d.h.

This is the only place the Python host side touches native code.  There is NO fallback: if the
shared library is missing and cannot be built, importing the rasterizer raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build

_LIB = None


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p)]


class Geometry(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_degree", C.c_int32), ("M", C.c_int32), ("C", C.c_int32), ("use_shs", C.c_int32),
                ("gamma", C.c_float), ("scale_modifier", C.c_float), ("background_depth", C.c_float),
                ("background", C.c_void_p), ("vertex", C.c_void_p), ("shs", C.c_void_p), ("feature", C.c_void_p),
                ("opacity", C.c_void_p), ("model", C.c_void_p)]  # model: POINTER(ModelInputs) or NULL


class ModelInputs(C.Structure):
    """ts2d_model_inputs: parameter-space inputs (VanillaTS_model.py:608-647 done inside the per-triangle kernels)."""
    _fields_ = [("f_dc", C.c_void_p), ("f_rest", C.c_void_p), ("opacity_logit", C.c_void_p), ("ste_threshold", C.c_float),
                ("rescale_ratio", C.c_float), ("bg_depth_from_vertices", C.c_int32)]


class ModelGrads(C.Structure):
    """ts2d_model_grads: gradients w.r.t. the raw parameters + the in-place training statistics (VanillaTS_model.py:347-363)."""
    _fields_ = [("dL_df_dc", C.c_void_p), ("dL_df_rest", C.c_void_p), ("gradient_accum", C.c_void_p), ("gradient_denom", C.c_void_p),
                ("contrib_sum", C.c_void_p), ("contrib_max", C.c_void_p), ("contrib_denom", C.c_void_p), ("max_radii2D", C.c_void_p),
                ("fwd_contrib_sum", C.c_void_p), ("fwd_contrib_max", C.c_void_p), ("radii_div", C.c_int32)]


class Flags(C.Structure):
    _fields_ = [("back_culling", C.c_int32), ("rich_info", C.c_int32), ("debug", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_world", C.c_int32), ("exact", C.c_int32), ("primitive", C.c_int32)]


MAX_RANKS = 8  # TS2D_MAX_RANKS


EXCHANGE_ADD_F32, EXCHANGE_MAX_U32 = 0, 1  # TS2D_EXCHANGE_*


class FrameCounters(C.Structure):
    """ts2d_frame_counters: device-side counters of a frame (R, rows of the atomics-free gradient write-back)."""
    _fields_ = [("num_rendered", C.c_int64), ("backward_rows", C.c_int64)]


class ForwardOut(C.Structure):
    _fields_ = [("out_feature", C.c_void_p), ("radii", C.c_void_p), ("depth", C.c_void_p), ("normal", C.c_void_p),
                ("contrib_sum", C.c_void_p), ("contrib_max", C.c_void_p)]


class LossIn(C.Structure):
    _fields_ = [("dL_dout_feature", C.c_void_p), ("dL_dout_depth", C.c_void_p), ("dL_dout_normal", C.c_void_p)]


class BackwardOut(C.Structure):
    _fields_ = [("dL_dvertex", C.c_void_p), ("dL_dcenter2D", C.c_void_p), ("dL_dshs", C.c_void_p), ("dL_dfeature", C.c_void_p),
                ("dL_dopacity", C.c_void_p), ("model", C.c_void_p)]  # model: POINTER(ModelGrads) or NULL


# every symbol include/ts2d.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = {
    "ts2d_abi_version": (C.c_int, []),
    "ts2d_error_string": (C.c_char_p, [C.c_int]),
    "ts2d_geometry_state_bytes": (C.c_size_t, [C.c_int32]),
    "ts2d_binning_state_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "ts2d_image_state_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "ts2d_binning_capacity": (C.c_int64, [C.c_size_t]),
    "ts2d_backward_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_size_t, C.c_int64]),
    "ts2d_counters_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ts2d_counters_destroy": (None, [C.c_void_p]),
    "ts2d_counters_num_rendered": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ts2d_counters_backward_rows": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ts2d_read_counters": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(FrameCounters), C.c_void_p]),
    "ts2d_forward": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                               C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(ForwardOut), C.c_void_p, C.c_void_p]),
    "ts2d_forward_geometry": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_int64), C.c_void_p]),
    "ts2d_forward_render": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(ForwardOut), C.c_void_p, C.c_void_p]),
    "ts2d_backward": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                C.c_void_p, C.POINTER(LossIn), C.POINTER(BackwardOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ts2d_backward_composite": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p, C.POINTER(LossIn), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ts2d_backward_geometry": (C.c_int, [C.POINTER(Camera), C.POINTER(Geometry), C.POINTER(Flags), C.c_void_p, C.c_void_p,
                                         C.POINTER(BackwardOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ts2d_export_geometry": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 10 + [C.c_void_p]),
    "ts2d_export_geometry3d": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_void_p]),
    "ts2d_export_model": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ts2d_exchange_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ts2d_exchange_allreduce": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "ts2d_image_loss_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "ts2d_image_loss_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "ts2d_image_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_void_p]),
    "ts2d_depth_normal_loss_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "ts2d_depth_normal_loss_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_void_p,
                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "ts2d_depth_normal_loss_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                  C.c_void_p, C.c_void_p]),
    "ts2d_downsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ts2d_downsample_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ts2d_export_binning": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "ts2d_export_image": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ts2d_profile_enable": (C.c_int, [C.c_int]),
    "ts2d_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

ABI_VERSION = 7  # TS2D_ABI_VERSION (include/ts2d.h)
PRIMITIVES = {"2D": 0, "3D": 1}  # TS2D_PRIMITIVE_* (include/ts2d.h)
STAGES = ("preprocess", "order_scan", "binning", "render_fwd", "render_bwd", "preprocess_bwd", "bwd_prepare", "bwd_reduce")


def lib_path() -> Path:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the sources are newer / the .so is absent and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if os.environ.get("TS2D_NO_AUTOBUILD", "0") != "1" and _build.needs_build():
        try:
            _build.build()
        except FileNotFoundError as ex:  # no nvcc on this box: an existing .so is used as it is (its ABI version is checked below)
            if not path.exists():
                raise RuntimeError(f"libts2d.so is missing and could not be built: {ex}") from ex
        # a compile / link error with newer sources is NOT swallowed: running a stale library against current struct layouts
        # would corrupt memory silently
    if not path.exists():
        raise RuntimeError(f"{path} not found: run `python -m triangle_splatting_b200.build` (needs nvcc); there is no CPU fallback")
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    have = int(lib.ts2d_abi_version())
    if have != ABI_VERSION:
        raise RuntimeError(f"{path} has ABI version {have}, this package binds version {ABI_VERSION}: rebuild it "
                           "(python -m triangle_splatting_b200.build --force)")
    _LIB = lib
    return lib


def error_string(code: int) -> str:
    return load().ts2d_error_string(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what}: {error_string(code)} (ts2d code {code})")
