"""Host-side mirror of the reference's pybind module ``diff_triangle_rasterization_2D._C``.

Same two entry points, argument order, return tuples and error behaviour as
R2D/ext.cpp:4-9 / R2D/src/extension_interface.cu:19-260, implemented on top of the C ABI of
libts2d.so (include/ts2d.h).  torch is used only for device memory (outputs + the three opaque
state tensors), the current CUDA stream and the device guard.

Extra keyword-only arguments (no counterpart in the reference):
    shard=(rank, world)  image-space tile sharding for multi-GPU (see distributed.py)
    primitive="2D"|"3D"  which of the reference's two rasterizer packages the call stands in for: "2D" =
                         diff_triangle_rasterization_2D (screen-space triangles, the north-star path), "3D" =
                         diff_triangle_rasterization_3D (ray / plane intersection in view space; identical pybind
                         signatures, R3D/ext.cpp:4-9 / R3D/src/extension_interface.cu:19-246)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import torch

from . import _lib

MAX_CHANNELS = 3  # R2D/src/config.h:3

# Arithmetic mode of the composite kernels (ts2d_flags.exact): False = fast path with exact re-evaluation inside the
# decision bands (default), True = op-for-op mirror of the reference's per-pair arithmetic.  TS2D_EXACT=1 forces the mirror.
EXACT = os.environ.get("TS2D_EXACT", "0") == "1"


def set_exact(flag: bool) -> bool:
    """Select the arithmetic mode for subsequent calls; returns the previous setting."""
    global EXACT
    old, EXACT = EXACT, bool(flag)
    return old


def _ptr(t: torch.Tensor | None):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _derive(vertex, shs, feature):
    """P, use_shs, C, M exactly as extension_interface.cu:41-50."""
    P = vertex.size(0) if vertex.dim() > 0 else 0
    use_shs = feature.dim() <= 1 or (feature.size(0) == 0 and shs.size(0) > 0)
    Cn = 3 if use_shs else feature.size(1)
    M = 0
    if shs.dim() > 0 and shs.size(0) != 0:
        M = shs.size(1)
    return P, bool(use_shs), int(Cn), int(M)


def _check_inputs(vertex, shs, feature, background, gamma, use_shs, Cn):
    # extension_interface.cu:53-76 (AT_ERROR -> RuntimeError)
    if vertex.dim() != 3 or vertex.size(1) != 3 or vertex.size(2) != 3:
        raise RuntimeError("vertex must have dimensions (num_points, 3, 3)")
    if not use_shs and feature.dim() != 2:
        raise RuntimeError("feature must have dimensions (num_points, num_channels)")
    if use_shs and shs.dim() != 3:
        raise RuntimeError("shs must have dimensions (num_points, (1 + sh_degree) ** 2, 3)")
    if Cn > MAX_CHANNELS:
        raise RuntimeError("feature's num_channels can't be larger than MAX_CHANNELS")
    if Cn != background.size(0):
        raise RuntimeError("background must have the same number of channels as feature")
    if gamma < 0.0:
        raise RuntimeError("gamma must be larger than 0")


def _require_cuda_f32(name, t):
    if t.numel() == 0:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this rasterizer has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")


def _structs(image_width, image_height, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
             background_depth, background, vertex, shs, feature, opacity, P, use_shs, Cn, M, back_culling, rich_info, debug, shard,
             primitive="2D"):
    cam = _lib.Camera(int(image_width), int(image_height), float(tan_fovx), float(tan_fovy), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos))
    geom = _lib.Geometry(int(P), int(sh_degree), int(M), int(Cn), int(use_shs), float(gamma), float(scale_modifier), float(background_depth),
                         _ptr(background), _ptr(vertex), _ptr(shs) if use_shs else None, None if use_shs else _ptr(feature), _ptr(opacity))
    flags = _lib.Flags(int(bool(back_culling)), int(bool(rich_info)), int(bool(debug)), int(shard[0]), int(shard[1]), int(EXACT),
                       _lib.PRIMITIVES[primitive])
    return cam, geom, flags


def rasterize_triangles(image_width: int, image_height: int, tan_fovx: float, tan_fovy: float, viewmatrix: torch.Tensor,
                        projmatrix: torch.Tensor, campos: torch.Tensor, sh_degree: int, gamma: float, scale_modifier: float,
                        background_depth: float, background: torch.Tensor, vertex: torch.Tensor, shs: torch.Tensor, feature: torch.Tensor,
                        opacity: torch.Tensor, back_culling: bool, rich_info: bool, debug: bool, *, shard: Tuple[int, int] = (0, 1),
                        primitive: str = "2D"):
    """-> (num_rendered:int, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer)

    Mirrors rasterizeTrianglesForward (extension_interface.cu:19-152)."""
    lib = _lib.load()
    P, use_shs, Cn, M = _derive(vertex, shs, feature)
    H, W = int(image_height), int(image_width)
    _check_inputs(vertex, shs, feature, background, gamma, use_shs, Cn)
    tensors = dict(viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos, background=background, vertex=vertex, shs=shs,
                   feature=feature, opacity=opacity)
    if not all(t.is_contiguous() for t in tensors.values()):
        raise RuntimeError("input tensors must be contiguous")
    for k, t in tensors.items():
        _require_cuda_f32(k, t)
    if use_shs and P > 0 and (sh_degree < 0 or sh_degree > 3 or (sh_degree + 1) ** 2 > M):
        raise RuntimeError(_lib.error_string(-9))
    if P > 0 and opacity.numel() != P:
        raise RuntimeError("opacity must have dimensions (num_points, 1)")

    dev = vertex.device if vertex.is_cuda else background.device
    f32 = dict(device=dev, dtype=torch.float32)
    u8 = dict(device=dev, dtype=torch.uint8)
    sharded = shard[1] > 1
    alloc = torch.zeros if (P == 0 or sharded) else torch.empty  # every owned pixel is written by the composite kernel
    out_feature = alloc((Cn, H, W), **f32)
    radii = torch.zeros((P,), device=dev, dtype=torch.int32) if P == 0 else torch.empty((P,), device=dev, dtype=torch.int32)
    if rich_info:
        depth = alloc((H, W), **f32)
        normal = alloc((3, H, W), **f32)
        contrib_sum = torch.empty((P,), **f32) if P else torch.zeros((0,), **f32)
        contrib_max = torch.empty((P,), **f32) if P else torch.zeros((0,), **f32)
    else:
        depth = torch.empty((0,), **f32)
        normal = torch.empty((0,), **f32)
        contrib_sum = torch.empty((0,), **f32)
        contrib_max = torch.empty((0,), **f32)
    if P == 0:  # extension_interface.cu:130 -- zero outputs, empty state
        e = torch.empty((0,), **u8)
        return 0, out_feature, radii, depth, normal, contrib_sum, contrib_max, e, e.clone(), e.clone()

    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        cam, geom, flags = _structs(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
                                    background_depth, background, vertex, shs, feature, opacity, P, use_shs, Cn, M, back_culling, rich_info,
                                    debug, shard, primitive)
        gbytes = lib.ts2d_geometry_state_bytes(P)
        geometryBuffer = torch.empty((gbytes,), **u8)
        num_rendered = C.c_int64(0)
        _lib.check(lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer), gbytes,
                                             C.byref(num_rendered), stream), "ts2d_forward_geometry")
        R = int(num_rendered.value)
        bbytes = lib.ts2d_binning_state_bytes(R, W, H)
        ibytes = lib.ts2d_image_state_bytes(W, H)
        binningBuffer = torch.empty((bbytes,), **u8)
        imageBuffer = torch.empty((ibytes,), **u8)
        out = _lib.ForwardOut(_ptr(out_feature), _ptr(radii), _ptr(depth), _ptr(normal), _ptr(contrib_sum), _ptr(contrib_max))
        _lib.check(lib.ts2d_forward_render(C.byref(cam), C.byref(geom), C.byref(flags), R, _ptr(geometryBuffer), _ptr(binningBuffer), bbytes,
                                           _ptr(imageBuffer), ibytes, C.byref(out), stream), "ts2d_forward_render")
    return R, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer


def rasterize_triangles_backward(tan_fovx: float, tan_fovy: float, viewmatrix: torch.Tensor, projmatrix: torch.Tensor, campos: torch.Tensor,
                                 sh_degree: int, gamma: float, scale_modifier: float, background_depth: float, background: torch.Tensor,
                                 vertex: torch.Tensor, shs: torch.Tensor, feature: torch.Tensor, opacity: torch.Tensor, num_rendered: int,
                                 radii: torch.Tensor, geometryBuffer: torch.Tensor, binningBuffer: torch.Tensor, imageBuffer: torch.Tensor,
                                 dL_dout_feature: torch.Tensor, dL_dout_depth: torch.Tensor | None, dL_dout_normal: torch.Tensor | None,
                                 rich_info: bool, debug: bool, *, shard: Tuple[int, int] = (0, 1), primitive: str = "2D"):
    """-> (dL_dvertex (P,3,3), dL_dcenter2D (P,2), dL_dshs (P,M,3), dL_dfeature (P,C), dL_dopacity (P,1))

    Mirrors rasterizeTrianglesBackward (extension_interface.cu:154-260).  Unlike the reference's
    Python wrapper (which raises UnboundLocalError, __init__.py:114-142) rich_info=False is accepted:
    pass None for dL_dout_depth / dL_dout_normal."""
    lib = _lib.load()
    P, use_shs, Cn, M = _derive(vertex, shs, feature)
    H, W = int(dL_dout_feature.size(1)), int(dL_dout_feature.size(2))
    ts = [viewmatrix, projmatrix, campos, background, vertex, shs, feature, opacity, radii, geometryBuffer, binningBuffer, imageBuffer,
          dL_dout_feature]
    if rich_info:
        ts += [dL_dout_depth, dL_dout_normal]
    if not all(t.is_contiguous() for t in ts):
        raise RuntimeError("input tensors must be contiguous")
    dev = vertex.device
    f32 = dict(device=dev, dtype=torch.float32)
    if P == 0:
        return (torch.zeros((P, 3, 3), **f32), torch.zeros((P, 2), **f32), torch.zeros((P, M, 3), **f32), torch.zeros((P, Cn), **f32),
                torch.zeros((P, 1), **f32))
    _require_cuda_f32("dL_dout_feature", dL_dout_feature)
    dL_dvertex = torch.empty((P, 3, 3), **f32)
    dL_dcenter2D = torch.empty((P, 2), **f32)
    dL_dshs = torch.empty((P, M, 3), **f32)
    dL_dfeature = torch.empty((P, Cn), **f32)
    dL_dopacity = torch.empty((P, 1), **f32)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        cam, geom, flags = _structs(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
                                    background_depth, background, vertex, shs, feature, opacity, P, use_shs, Cn, M, False, rich_info, debug,
                                    shard, primitive)
        sbytes = lib.ts2d_backward_scratch_bytes(P)
        scratch = torch.empty((sbytes,), device=dev, dtype=torch.uint8)
        loss = _lib.LossIn(_ptr(dL_dout_feature), _ptr(dL_dout_depth) if rich_info else None, _ptr(dL_dout_normal) if rich_info else None)
        out = _lib.BackwardOut(_ptr(dL_dvertex), _ptr(dL_dcenter2D), _ptr(dL_dshs), _ptr(dL_dfeature), _ptr(dL_dopacity))
        if shard[1] > 1:
            # tile-sharded: composite over this rank's tiles, sum the 64 B/triangle accumulators over the ranks (NCCL on the
            # current stream), then the per-triangle stage runs replicated on identical data -> identical gradients everywhere
            from . import distributed

            _lib.check(lib.ts2d_backward_composite(C.byref(cam), C.byref(geom), C.byref(flags), int(num_rendered), _ptr(geometryBuffer),
                                                   _ptr(binningBuffer), _ptr(imageBuffer), C.byref(loss), _ptr(scratch), sbytes, stream),
                       "ts2d_backward_composite")
            distributed.reduce_accumulators(scratch.view(torch.float32))
            _lib.check(lib.ts2d_backward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer),
                                                  C.byref(out), _ptr(scratch), sbytes, stream), "ts2d_backward_geometry")
        else:
            _lib.check(lib.ts2d_backward(C.byref(cam), C.byref(geom), C.byref(flags), int(num_rendered), _ptr(radii), _ptr(geometryBuffer),
                                         _ptr(binningBuffer), _ptr(imageBuffer), C.byref(loss), C.byref(out), _ptr(scratch), sbytes, stream),
                       "ts2d_backward")
    return dL_dvertex, dL_dcenter2D, dL_dshs, dL_dfeature, dL_dopacity
