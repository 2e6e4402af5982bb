"""Host-side mirror of the reference's pybind module ``diff_triangle_rasterization_2D._C``.

Same two entry points, argument order, return tuples and error behaviour as
R2D/ext.cpp:4-9 / R2D/src/extension_interface.cu:19-260, implemented on top of the C ABI of
libts2d.so (include/ts2d.h).  torch is used only for device memory (outputs + the three opaque
state tensors), the current CUDA stream and the device guard.

Extra keyword-only arguments (no counterpart in the reference):
    shard=(rank, world)  image-space tile sharding for multi-GPU (see distributed.py)
    model=ModelInputs    parameter-space inputs: raw _f_dc / _f_rest / _opacity logits, STE
                         threshold, gamma_rescale ratio, device-side background depth -- the Python preamble of
                         src/diff_recon/models/VanillaTS_model.py:608-647 done inside the per-triangle kernels; `shs` and
                         `opacity` are then passed as None.  `stats=` (backward) adds the in-place training statistics
                         of VanillaTS_model.py:347-363 (rank 3)
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


# One-enqueue forward (ts2d_forward): the binning state is sized for a capacity derived from the largest instance count seen so
# far for the same problem shape, every kernel takes the true count R from device memory, and R is only read back -- behind an
# event recorded right after the scan, while the rest of the frame is already queued -- to return it and to detect an overflow.
# TS2D_SYNC_FORWARD=1 forces the two-call path (ts2d_forward_geometry blocks for R, like the reference's cudaMemcpy,
# rasterizer.cu:190-193); it is also what the first frame of a shape and debug=True use.
SYNC_FORWARD = os.environ.get("TS2D_SYNC_FORWARD", "0") == "1"
_R_SEEN = {}  # (device index, P, W, H, primitive, shard) -> largest R seen
CAPACITY_MARGIN = 1 << 20  # instances added on top of the largest R seen (16 B each: memory is not what limits this path)
# Backward, same idea: the row array of the atomics-free gradient write-back is sized from the largest row count seen for the
# shape, the whole backward pass is enqueued, and only then the frame's true row count is read (it landed with the end of the
# forward pass, so the wait never stalls the queue); the kernels never store past the capacity, a frame that needed more rows is
# composited again with the exact size.  TS2D_SYNC_BACKWARD=1 waits for the count first.
SYNC_BACKWARD = os.environ.get("TS2D_SYNC_BACKWARD", "0") == "1"
_ROWS_SEEN = {}  # (device index, P, W, H, primitive, shard) -> largest backward row count seen
ROWS_MARGIN = 1 << 16  # rows (64 B each) on top of 1.25 x the largest count seen

# Host layer of the reference-shaped single-GPU call: the compiled pybind11 module _C_native (csrc/ts2d_pybind.cpp -- what the
# reference's ext.cpp / extension_interface.cu are to its library) when it is built, else the ctypes code below; both sit on the same
# C ABI and the same libts2d.so.  TS2D_HOST=ctypes forces the Python layer, TS2D_HOST=native makes a missing module an error.
# The tile-sharded and the parameter-space calls always take the ctypes layer.
HOST = os.environ.get("TS2D_HOST", "auto")
_NATIVE = None


def native():
    """The compiled host layer, or None."""
    global _NATIVE
    if _NATIVE is None:
        _NATIVE = False
        if HOST != "ctypes":
            try:
                from . import build as _build

                _lib.load()
                if os.environ.get("TS2D_NO_AUTOBUILD", "0") != "1" and _build.native_needs_build():
                    _build.build_native()
                from . import _C_native

                if _C_native.abi_version() != _lib.ABI_VERSION:
                    raise RuntimeError(f"_C_native was built against ABI {_C_native.abi_version()}, this package binds {_lib.ABI_VERSION}")
                _C_native.configure(SYNC_FORWARD, SYNC_BACKWARD, CAPACITY_MARGIN, ROWS_MARGIN)
                _NATIVE = _C_native
            except Exception as ex:  # noqa: BLE001  -- the ctypes layer does the same job (same library, same kernels)
                if HOST == "native":
                    raise
                import warnings

                warnings.warn(f"triangle_splatting_b200: compiled host layer _C_native unavailable ({ex}); using the ctypes layer")
    return _NATIVE or None


def configure(sync_forward=None, sync_backward=None, capacity_margin=None, rows_margin=None):
    """Knobs of the one-enqueue paths, for both host layers; returns the previous (sync_forward, sync_backward, capacity_margin, rows_margin)."""
    global SYNC_FORWARD, SYNC_BACKWARD, CAPACITY_MARGIN, ROWS_MARGIN
    old = (SYNC_FORWARD, SYNC_BACKWARD, CAPACITY_MARGIN, ROWS_MARGIN)
    SYNC_FORWARD = SYNC_FORWARD if sync_forward is None else bool(sync_forward)
    SYNC_BACKWARD = SYNC_BACKWARD if sync_backward is None else bool(sync_backward)
    CAPACITY_MARGIN = CAPACITY_MARGIN if capacity_margin is None else int(capacity_margin)
    ROWS_MARGIN = ROWS_MARGIN if rows_margin is None else int(rows_margin)
    if native() is not None:
        native().configure(SYNC_FORWARD, SYNC_BACKWARD, CAPACITY_MARGIN, ROWS_MARGIN)
    return old


def forget_shapes():
    """Drop the instance / row counts remembered per problem shape (the next frame of every shape takes the synchronous path)."""
    _R_SEEN.clear()
    _ROWS_SEEN.clear()
    if native() is not None:
        native().forget_shapes()


def poison_shapes(r: int = 1, rows: int = 1):
    """Tests: pretend every shape seen so far had `r` instances and `rows` backward rows (forces the overflow repairs)."""
    for k in _R_SEEN:
        _R_SEEN[k] = r
    for k in _ROWS_SEEN:
        _ROWS_SEEN[k] = rows
    if native() is not None:
        native().poison_shapes(r, rows)


def shapes_seen():
    """-> (largest R remembered, largest backward row count remembered) over both host layers."""
    r, rows = max(_R_SEEN.values(), default=0), max(_ROWS_SEEN.values(), default=0)
    if native() is not None:
        _, nr, _, nrows = native().shapes_seen()
        r, rows = max(r, nr), max(rows, nrows)
    return r, rows


def _capacity_for(key):
    r = _R_SEEN.get(key)
    if r is None or SYNC_FORWARD:
        return None
    return max(int(1.5 * r), r + CAPACITY_MARGIN)


class FrameCounters:
    """ts2d_counters handle: pinned host memory + events for the device-side counters of one forward pass."""

    _free = []
    created = 0  # handles made so far (a steady-state training loop reuses the ones it gave back)

    def __init__(self):
        lib = _lib.load()
        if FrameCounters._free:
            self.h = FrameCounters._free.pop()
        else:
            h = C.c_void_p()
            _lib.check(lib.ts2d_counters_create(C.byref(h)), "ts2d_counters_create")
            FrameCounters.created += 1
            self.h = h

    def num_rendered(self) -> int:
        v = C.c_int64(0)
        _lib.check(_lib.load().ts2d_counters_num_rendered(self.h, C.byref(v)), "ts2d_counters_num_rendered")
        return int(v.value)

    def backward_rows(self) -> int:
        v = C.c_int64(0)
        _lib.check(_lib.load().ts2d_counters_backward_rows(self.h, C.byref(v)), "ts2d_counters_backward_rows")
        return int(v.value)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h is not None and FrameCounters is not None:
            FrameCounters._free.append(h)  # events and pinned memory are reused by a later frame


class NumRendered(int):
    """The `num_rendered` the reference returns (an int) that also carries the frame's counters handle from forward to backward:
    the reference-shaped wrappers pass it through unchanged (ctx.num_rendered, __init__.py:93,123)."""

    def __new__(cls, value, counters=None):
        obj = super().__new__(cls, value)
        obj.counters = counters
        return obj


class ModelInputs:
    """Host-side mirror of ts2d_model_inputs (include/ts2d.h)."""

    def __init__(self, f_dc: torch.Tensor, f_rest: torch.Tensor | None, opacity_logit: torch.Tensor, ste_threshold: float | None = None,
                 rescale_ratio: float = 1.0, bg_depth_from_vertices: bool = True):
        self.f_dc, self.f_rest, self.opacity_logit = f_dc, f_rest, opacity_logit
        self.ste_threshold = -1.0 if ste_threshold is None else float(ste_threshold)
        self.rescale_ratio = float(rescale_ratio)
        self.bg_depth_from_vertices = bool(bg_depth_from_vertices)

    @property
    def M(self) -> int:
        return 1 + (0 if self.f_rest is None or self.f_rest.numel() == 0 else int(self.f_rest.size(1)))

    def check(self, P: int):
        if self.f_dc.dim() != 3 or tuple(self.f_dc.shape) != (P, 1, 3):
            raise RuntimeError("f_dc must have dimensions (num_points, 1, 3)")
        if self.M > 1 and (self.f_rest.dim() != 3 or self.f_rest.size(0) != P or self.f_rest.size(2) != 3):
            raise RuntimeError("f_rest must have dimensions (num_points, (1 + sh_degree) ** 2 - 1, 3)")
        if self.opacity_logit.numel() != P:
            raise RuntimeError("opacity must have dimensions (num_points, 1)")
        if not self.rescale_ratio > 0.0:
            raise RuntimeError("rescale_ratio must be positive")
        ts = [self.f_dc, self.opacity_logit] + ([self.f_rest] if self.M > 1 else [])
        if not all(t.is_contiguous() for t in ts):
            raise RuntimeError("input tensors must be contiguous")
        for t in ts:
            _require_cuda_f32("model input", t)

    def struct(self):
        return _lib.ModelInputs(_ptr(self.f_dc), _ptr(self.f_rest) if self.M > 1 else None, _ptr(self.opacity_logit), self.ste_threshold,
                                self.rescale_ratio, int(self.bg_depth_from_vertices))


STAT_FIELDS = ("gradient_accum", "gradient_denom", "contrib_sum", "contrib_max", "contrib_denom", "max_radii2D")


def _fabric_for(shard, dev):
    """NVLink peer-memory fabric for a tile-sharded call, or None (then NCCL collectives above this module do the exchange)."""
    if shard[1] <= 1:
        return None
    from . import distributed

    return distributed.fabric(dev)


def _home_slice(n4: int, rank: int, world: int):
    """This rank's slice (first, count) of an array of n4 (a multiple of 4) four-byte elements, both multiples of 4."""
    chunk = ((n4 + world - 1) // world + 3) // 4 * 4
    first = min(rank * chunk, n4)
    return first, min(chunk, n4 - first)


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
             primitive="2D", model=None):
    """-> (camera, geometry, flags, keepalive): `keepalive` holds the ctypes objects the structs point to."""
    mstruct = model.struct() if model is not None else None
    cam = _lib.Camera(int(image_width), int(image_height), float(tan_fovx), float(tan_fovy), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos))
    geom = _lib.Geometry(int(P), int(sh_degree), int(M), int(Cn), int(use_shs), float(gamma), float(scale_modifier), float(background_depth),
                         _ptr(background), _ptr(vertex), _ptr(shs) if (use_shs and model is None) else None,
                         None if use_shs else _ptr(feature), _ptr(opacity) if model is None else None,
                         C.cast(C.pointer(mstruct), C.c_void_p) if mstruct is not None else None)
    flags = _lib.Flags(int(bool(back_culling)), int(bool(rich_info)), int(bool(debug)), int(shard[0]), int(shard[1]), int(EXACT),
                       _lib.PRIMITIVES[primitive])
    return cam, geom, flags, mstruct


def rasterize_triangles(image_width: int, image_height: int, tan_fovx: float, tan_fovy: float, viewmatrix: torch.Tensor,
                        projmatrix: torch.Tensor, campos: torch.Tensor, sh_degree: int, gamma: float, scale_modifier: float,
                        background_depth: float, background: torch.Tensor, vertex: torch.Tensor, shs: torch.Tensor, feature: torch.Tensor,
                        opacity: torch.Tensor, back_culling: bool, rich_info: bool, debug: bool, *, shard: Tuple[int, int] = (0, 1),
                        primitive: str = "2D", model: ModelInputs | None = None):
    """-> (num_rendered:int, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer)

    Mirrors rasterizeTrianglesForward (extension_interface.cu:19-152)."""
    nat = native() if (model is None and shard[1] <= 1) else None
    if nat is not None:
        out = nat.rasterize_triangles(image_width, image_height, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
                                      background_depth, background, vertex, shs, feature, opacity, back_culling, rich_info, debug,
                                      _lib.PRIMITIVES[primitive], EXACT)
        return (NumRendered(out[0], out[10]),) + out[1:10]
    lib = _lib.load()
    if model is not None:  # parameter-space inputs: shs / opacity come from the model's raw tensors
        shs = feature = opacity = torch.empty(0)
        P, use_shs, Cn, M = (vertex.size(0) if vertex.dim() > 0 else 0), True, 3, model.M
        if vertex.dim() != 3 or vertex.size(1) != 3 or vertex.size(2) != 3:
            raise RuntimeError("vertex must have dimensions (num_points, 3, 3)")
        if Cn != background.size(0):
            raise RuntimeError("background must have the same number of channels as feature")
        if gamma < 0.0:
            raise RuntimeError("gamma must be larger than 0")
        if P > 0:
            model.check(P)
    else:
        P, use_shs, Cn, M = _derive(vertex, shs, feature)
        _check_inputs(vertex, shs, feature, background, gamma, use_shs, Cn)
    H, W = int(image_height), int(image_width)
    tensors = dict(viewmatrix=viewmatrix, projmatrix=projmatrix, campos=campos, background=background, vertex=vertex, shs=shs,
                   feature=feature, opacity=opacity)
    if not all(t.is_contiguous() for t in tensors.values()):
        raise RuntimeError("input tensors must be contiguous")
    for k, t in tensors.items():
        _require_cuda_f32(k, t)
    if use_shs and P > 0 and (sh_degree < 0 or sh_degree > 3 or (sh_degree + 1) ** 2 > M):
        raise RuntimeError(_lib.error_string(-9))
    if P > 0 and model is None and opacity.numel() != P:
        raise RuntimeError("opacity must have dimensions (num_points, 1)")

    dev = vertex.device if vertex.is_cuda else background.device
    f32 = dict(device=dev, dtype=torch.float32)
    u8 = dict(device=dev, dtype=torch.uint8)
    sharded = shard[1] > 1
    alloc = torch.zeros if (P == 0 or sharded) else torch.empty  # every owned pixel is written by the composite kernel
    radii = torch.zeros((P,), device=dev, dtype=torch.int32) if P == 0 else torch.empty((P,), device=dev, dtype=torch.int32)
    fab = _fabric_for(shard, dev) if (sharded and P > 0) else None
    if fab is not None:
        # NVLink peer memory (include/ts2d.h: "Multi-GPU exchange"): the outputs live in this rank's replica of ONE symmetric buffer
        # [image | depth | normal | pad | contrib_sum | contrib_max]; the composite kernel fills the owned tiles and this rank's partial
        # contrib statistics, the exchange kernels behind it complete every replica -- no collective, no zero-filled planes
        n_img, n_pix = Cn * H * W, H * W
        pad4 = lambda n: (n + 3) // 4 * 4
        n_planes = Cn + 4 if rich_info else Cn
        o_sum = pad4(n_planes * n_pix)
        o_max = o_sum + (pad4(P) if rich_info else 0)
        frame, mc_base, fab_h = fab.buffer("frame", o_max + (pad4(P) if rich_info else 0))
        out_feature = frame[:n_img].view(Cn, H, W)
        if rich_info:
            depth, normal = frame[n_img:n_img + n_pix].view(H, W), frame[n_img + n_pix:n_img + 4 * n_pix].view(3, H, W)
            contrib_sum, contrib_max = frame[o_sum:o_sum + P], frame[o_max:o_max + P]
        else:
            depth = normal = contrib_sum = contrib_max = torch.empty((0,), **f32)
    elif sharded and rich_info and P > 0:
        # one buffer [image | depth | normal | contrib_sum]: the ranks' partial frames are summed with ONE in-place all-reduce
        n_img, n_pix = Cn * H * W, H * W
        flat = torch.zeros((n_img + n_pix + 3 * n_pix + P,), **f32)
        out_feature = flat[:n_img].view(Cn, H, W)
        depth = flat[n_img:n_img + n_pix].view(H, W)
        normal = flat[n_img + n_pix:n_img + 4 * n_pix].view(3, H, W)
        contrib_sum = flat[n_img + 4 * n_pix:]
        contrib_max = torch.empty((P,), **f32)
    elif rich_info:
        out_feature = alloc((Cn, H, W), **f32)
        depth = alloc((H, W), **f32)
        normal = alloc((3, H, W), **f32)
        contrib_sum = torch.empty((P,), **f32) if P else torch.zeros((0,), **f32)
        contrib_max = torch.empty((P,), **f32) if P else torch.zeros((0,), **f32)
    else:
        out_feature = alloc((Cn, H, W), **f32)
        depth = torch.empty((0,), **f32)
        normal = torch.empty((0,), **f32)
        contrib_sum = torch.empty((0,), **f32)
        contrib_max = torch.empty((0,), **f32)
    if P == 0:  # extension_interface.cu:130 -- zero outputs, empty state
        e = torch.empty((0,), **u8)
        return 0, out_feature, radii, depth, normal, contrib_sum, contrib_max, e, e.clone(), e.clone()

    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        cam, geom, flags, _keep = _structs(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
                                           background_depth, background, vertex, shs, feature, opacity, P, use_shs, Cn, M, back_culling,
                                           rich_info, debug, shard, primitive, model)
        gbytes = lib.ts2d_geometry_state_bytes(P)
        ibytes = lib.ts2d_image_state_bytes(W, H)
        geometryBuffer = torch.empty((gbytes,), **u8)
        imageBuffer = torch.empty((ibytes,), **u8)
        counters = FrameCounters()
        shape_key = (dev.index, P, W, H, primitive, tuple(shard))
        cap = None if debug else _capacity_for(shape_key)
        out = _lib.ForwardOut(_ptr(out_feature), _ptr(radii), _ptr(depth), _ptr(normal), _ptr(contrib_sum), _ptr(contrib_max))
        R = None
        if cap is not None:
            # one enqueue, no synchronisation in front of the binning / composite kernels
            bbytes = lib.ts2d_binning_state_bytes(cap, W, H)
            binningBuffer = torch.empty((bbytes,), **u8)
            _lib.check(lib.ts2d_forward(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer), gbytes,
                                        _ptr(binningBuffer), bbytes, _ptr(imageBuffer), ibytes, C.byref(out), counters.h, stream), "ts2d_forward")
            R = counters.num_rendered()  # waits for the scan only: the GPU is busy with the rest of the frame
            if R > lib.ts2d_binning_capacity(bbytes):
                cap = None  # the capacity guess was too small: the frame is invalid, render again below on the same geometry state
        if cap is None:
            if R is None:
                num_rendered = C.c_int64(0)
                _lib.check(lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer), gbytes,
                                                     C.byref(num_rendered), stream), "ts2d_forward_geometry")
                R = int(num_rendered.value)
            bbytes = lib.ts2d_binning_state_bytes(R, W, H)
            binningBuffer = torch.empty((bbytes,), **u8)
            _lib.check(lib.ts2d_forward_render(C.byref(cam), C.byref(geom), C.byref(flags), R, _ptr(geometryBuffer), _ptr(binningBuffer), bbytes,
                                               _ptr(imageBuffer), ibytes, C.byref(out), counters.h, stream), "ts2d_forward_render")
        _R_SEEN[shape_key] = max(R, _R_SEEN.get(shape_key, 0))
        R = NumRendered(R, counters)
        if fab is not None:
            fab_h.barrier(channel=0)     # every rank's tiles and partial statistics are complete, and nobody still reads the previous frame
            _lib.check(lib.ts2d_exchange_tiles(C.c_void_p(frame.data_ptr()), C.c_void_p(mc_base), n_planes, W, H, shard[0], shard[1], stream),
                       "ts2d_exchange_tiles")
            if rich_info:
                first, count = _home_slice(pad4(P), shard[0], shard[1])
                _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(mc_base), o_sum + first, count, _lib.EXCHANGE_ADD_F32, stream), "ts2d_exchange_allreduce")
                _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(mc_base), o_max + first, count, _lib.EXCHANGE_MAX_U32, stream), "ts2d_exchange_allreduce")
            fab_h.barrier(channel=0)     # every rank has published
            mine = frame.clone()         # the symmetric buffer is reused by the next frame
            out_feature = mine[:n_img].view(Cn, H, W)
            if rich_info:
                depth, normal = mine[n_img:n_img + n_pix].view(H, W), mine[n_img + n_pix:n_img + 4 * n_pix].view(3, H, W)
                contrib_sum, contrib_max = mine[o_sum:o_sum + P], mine[o_max:o_max + P]
            out_feature._ts2d_assembled = True  # tells the autograd wrapper that no collective is needed
    return R, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer, imageBuffer


def rasterize_triangles_backward(tan_fovx: float, tan_fovy: float, viewmatrix: torch.Tensor, projmatrix: torch.Tensor, campos: torch.Tensor,
                                 sh_degree: int, gamma: float, scale_modifier: float, background_depth: float, background: torch.Tensor,
                                 vertex: torch.Tensor, shs: torch.Tensor, feature: torch.Tensor, opacity: torch.Tensor, num_rendered: int,
                                 radii: torch.Tensor, geometryBuffer: torch.Tensor, binningBuffer: torch.Tensor, imageBuffer: torch.Tensor,
                                 dL_dout_feature: torch.Tensor, dL_dout_depth: torch.Tensor | None, dL_dout_normal: torch.Tensor | None,
                                 rich_info: bool, debug: bool, *, shard: Tuple[int, int] = (0, 1), primitive: str = "2D",
                                 model: ModelInputs | None = None, stats: dict | None = None, fwd_contrib=None, radii_div: int = 1):
    """-> (dL_dvertex (P,3,3), dL_dcenter2D (P,2), dL_dshs (P,M,3), dL_dfeature (P,C), dL_dopacity (P,1))

    Mirrors rasterizeTrianglesBackward (extension_interface.cu:154-260).  Unlike the reference's
    Python wrapper (which raises UnboundLocalError, __init__.py:114-142) rich_info=False is accepted:
    pass None for dL_dout_depth / dL_dout_normal.

    With `model` (parameter-space inputs) the tuple is (dL_d_vertex, dL_dcenter2D, dL_d_f_dc (P,1,3), dL_d_f_rest (P,M-1,3),
    dL_d_opacity_logit (P,1)); `stats` (dict of the six (P,) tensors of VanillaTS_model.py:196-201, any subset) is updated in
    place for the visible triangles, `fwd_contrib` = (contrib_sum, contrib_max) of the forward pass."""
    nat = native() if (model is None and shard[1] <= 1 and stats is None) else None
    if nat is not None:
        ctr = getattr(num_rendered, "counters", None)
        if ctr is None or isinstance(ctr, nat.FrameCounters):
            return nat.rasterize_triangles_backward(tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier, background_depth,
                                                    background, vertex, shs, feature, opacity, int(num_rendered), radii, geometryBuffer, binningBuffer,
                                                    imageBuffer, dL_dout_feature, dL_dout_depth, dL_dout_normal, rich_info, debug, ctr,
                                                    _lib.PRIMITIVES[primitive], EXACT)
    lib = _lib.load()
    if model is not None:
        shs = feature = opacity = torch.empty(0, device=vertex.device)
        P, use_shs, Cn, M = vertex.size(0), True, 3, model.M
    else:
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
        if model is not None:
            return (torch.zeros((P, 3, 3), **f32), torch.zeros((P, 2), **f32), torch.zeros((P, 1, 3), **f32),
                    torch.zeros((P, M - 1, 3), **f32), torch.zeros((P, 1), **f32))
        return (torch.zeros((P, 3, 3), **f32), torch.zeros((P, 2), **f32), torch.zeros((P, M, 3), **f32), torch.zeros((P, Cn), **f32),
                torch.zeros((P, 1), **f32))
    _require_cuda_f32("dL_dout_feature", dL_dout_feature)
    dL_dvertex = torch.empty((P, 3, 3), **f32)
    dL_dcenter2D = torch.empty((P, 2), **f32)
    mgrads = None
    if model is not None:
        dL_df_dc = torch.empty((P, 1, 3), **f32)
        dL_df_rest = torch.empty((P, M - 1, 3), **f32)
        dL_dshs = None
        st = dict(stats or {})
        for k, t in st.items():
            if k not in STAT_FIELDS:
                raise RuntimeError(f"unknown statistics tensor {k!r}")
            if t.numel() != P or not t.is_contiguous():
                raise RuntimeError(f"statistics tensor {k} must be a contiguous (num_points,) tensor")
            _require_cuda_f32(k, t)
        if ("contrib_sum" in st or "contrib_max" in st) and fwd_contrib is None:
            raise RuntimeError("contrib statistics need the forward pass's contrib_sum / contrib_max")
        mgrads = _lib.ModelGrads(_ptr(dL_df_dc), _ptr(dL_df_rest), *[_ptr(st.get(k)) for k in STAT_FIELDS],
                                 _ptr(fwd_contrib[0]) if fwd_contrib is not None else None,
                                 _ptr(fwd_contrib[1]) if fwd_contrib is not None else None, int(radii_div))
    else:
        dL_dshs = torch.empty((P, M, 3), **f32)
    dL_dfeature = torch.empty((P, Cn), **f32)
    dL_dopacity = torch.empty((P, 1), **f32)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        cam, geom, flags, _keep = _structs(W, H, tan_fovx, tan_fovy, viewmatrix, projmatrix, campos, sh_degree, gamma, scale_modifier,
                                           background_depth, background, vertex, shs, feature, opacity, P, use_shs, Cn, M, False, rich_info,
                                           debug, shard, primitive, model)
        # rows of the atomics-free gradient write-back: from the counters the forward pass left behind (no wait in practice: that
        # forward finished long ago); a plain int (a caller that built the arguments itself) costs one blocking read
        ctr = getattr(num_rendered, "counters", None)
        rows_key = (dev.index, P, W, H, primitive, tuple(shard))
        rows = None  # the frame's row count, once known
        if EXACT or not (0.6 <= float(gamma) <= 64.0):
            rows = rows_cap = 0
        elif ctr is not None and not SYNC_BACKWARD and not debug and rows_key in _ROWS_SEEN:
            rows_cap = _ROWS_SEEN[rows_key] + _ROWS_SEEN[rows_key] // 4 + ROWS_MARGIN  # checked against the true count below
        elif ctr is not None:
            rows = rows_cap = ctr.backward_rows()
        else:
            fcn = _lib.FrameCounters()
            _lib.check(lib.ts2d_read_counters(_ptr(geometryBuffer), P, C.byref(fcn), stream), "ts2d_read_counters")
            rows = rows_cap = int(fcn.backward_rows)
        bbytes = binningBuffer.numel()
        loss = _lib.LossIn(_ptr(dL_dout_feature), _ptr(dL_dout_depth) if rich_info else None, _ptr(dL_dout_normal) if rich_info else None)
        out = _lib.BackwardOut(_ptr(dL_dvertex), _ptr(dL_dcenter2D), _ptr(dL_dshs), _ptr(dL_dfeature), _ptr(dL_dopacity),
                               C.cast(C.pointer(mgrads), C.c_void_p) if mgrads is not None else None)
        # tile-sharded over peer memory: the partial sums go straight into this rank's replica of a symmetric array
        fab = _fabric_for(shard, dev) if shard[1] > 1 else None
        acc_mc = acc_h = None
        if fab is not None:
            acc, acc_mc, acc_h = fab.buffer("accumulators", 16 * P)
        while True:
            sbytes = lib.ts2d_backward_scratch_bytes(P, bbytes, rows_cap)
            scratch = torch.empty((sbytes,), device=dev, dtype=torch.uint8)
            if rows is not None and shard[1] == 1:  # row count known up front: the whole pass in one call
                _lib.check(lib.ts2d_backward(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer), _ptr(binningBuffer),
                                             bbytes, _ptr(imageBuffer), C.byref(loss), C.byref(out), _ptr(scratch), sbytes, stream), "ts2d_backward")
                break
            # composite (K8 + row reduction) -> per-triangle sums: the first 16 P floats of the scratch, or the symmetric array
            if fab is None:
                acc = scratch[:64 * P].view(torch.float32)
            _lib.check(lib.ts2d_backward_composite(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(geometryBuffer), _ptr(binningBuffer), bbytes,
                                                   _ptr(imageBuffer), C.byref(loss), _ptr(scratch), sbytes, _ptr(acc) if fab is not None else None,
                                                   stream), "ts2d_backward_composite")
            if rows is None:
                rows = ctr.backward_rows()  # landed with the end of the forward pass: the composite above is already queued behind it
                if rows > rows_cap:
                    rows_cap = rows  # the guess was too small (rows beyond it were dropped): composite again, exact size
                    continue
            if shard[1] > 1:
                # sum the ranks' partial sums, then the per-triangle stage runs replicated on identical data -> identical gradients
                if fab is not None:
                    # every rank combines its slice of the triangles inside the switch and writes it back to all replicas
                    acc_h.barrier(channel=0)
                    first, count = _home_slice(16 * P, shard[0], shard[1])
                    _lib.check(lib.ts2d_exchange_allreduce(C.c_void_p(acc_mc), first, count, _lib.EXCHANGE_ADD_F32, stream), "ts2d_exchange_allreduce")
                    acc_h.barrier(channel=0)
                else:
                    from . import distributed

                    distributed.reduce_accumulators(acc)
            _lib.check(lib.ts2d_backward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), _ptr(radii), _ptr(geometryBuffer),
                                                  C.byref(out), _ptr(acc), 64 * P, stream), "ts2d_backward_geometry")
            break
        if rows_cap > 0:
            _ROWS_SEEN[rows_key] = max(rows, _ROWS_SEEN.get(rows_key, 0))
    if model is not None:
        return dL_dvertex, dL_dcenter2D, dL_df_dc, dL_df_rest, dL_dopacity
    return dL_dvertex, dL_dcenter2D, dL_dshs, dL_dfeature, dL_dopacity


def downsample(x: torch.Tensor, s: int, backward: bool = False) -> torch.Tensor:
    """render_up_scale epilogue (VanillaTS_model.py:647-655): F.interpolate(x, size=(H/s, W/s), mode="bilinear") of a planar
    (planes, H, W) image by an integer factor, or (backward=True) its adjoint applied to a (planes, H/s, W/s) gradient."""
    lib = _lib.load()
    _require_cuda_f32("image", x)
    if x.dim() != 3 or not x.is_contiguous():
        raise RuntimeError("downsample expects a contiguous (planes, H, W) tensor")
    planes, h, w = x.shape
    s = int(s)
    if not backward and (h % s or w % s):
        raise RuntimeError("image size must be a multiple of render_up_scale")
    oh, ow = (h * s, w * s) if backward else (h // s, w // s)
    out = torch.empty((planes, oh, ow), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        if backward:
            _lib.check(lib.ts2d_downsample_bwd(_ptr(x), _ptr(out), planes, w, h, s, stream), "ts2d_downsample_bwd")
        else:
            _lib.check(lib.ts2d_downsample(_ptr(x), _ptr(out), planes, ow, oh, s, stream), "ts2d_downsample")
    return out
