"""triangle_splatting_b200 -- B200-native differentiable 2D triangle-splatting rasterizer.

Drop-in for the reference package ``diff_triangle_rasterization_2D``
(R2D/diff_triangle_rasterization_2D/__init__.py): same ``TriangleRasterizationSettings`` (15 fields,
same order, :28-46), same ``TriangleRasterizer(nn.Module)`` (:167-187) and the same
``_RasterizeTriangles`` autograd.Function (:49-164), so ``diff_recon``'s
``TriangleRenderer`` (src/diff_recon/renderer/triangle_renderer.py:38-75) runs on it unchanged
(the top-level ``diff_triangle_rasterization_2D`` shim package re-exports these names).

The compute path is hand-written sm_100a CUDA behind a C ABI (include/ts2d.h, libts2d.so); there is
no CPU or PyTorch fallback.
"""
from __future__ import annotations

from typing import Callable, NamedTuple

import torch
import torch.nn as nn

from . import _C

__all__ = ["TriangleRasterizationSettings", "TriangleRasterizer", "_RasterizeTriangles", "TriangleRasterizer3D", "_RasterizeTriangles3D", "_C"]


def _cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def debug_run(func: Callable, *args, debug: bool = False, **kw):
    """Reference behaviour (__init__.py:14-25): with debug=True, snapshot the arguments on failure."""
    if not debug:
        return func(*args, **kw)
    func_name = func.__name__
    cpu_args = _cpu_deep_copy_tuple(args)
    try:
        return func(*args, **kw)
    except Exception as ex:
        torch.save(cpu_args, f"snapshot_{func_name}.dump")
        print(f"\nAn error occured in {func_name}. Writing snapshot_{func_name}.dump for debugging.")
        raise ex


class TriangleRasterizationSettings(NamedTuple):
    # Camera settings
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    # Geometry settings
    sh_degree: int
    gamma: float
    scale_modifier: float
    background_depth: float
    background: torch.Tensor
    # Rasterization settings
    back_culling: bool
    rich_info: bool
    debug: bool


def _shard():
    from . import distributed

    return distributed.current_shard()


class _RasterizeTriangles(torch.autograd.Function):
    PRIMITIVE = "2D"  # which reference package this Function stands in for (see _C.rasterize_triangles)

    @classmethod
    def apply(cls, *args, **kwargs):  # noqa: D102  -- tags the call with the primitive; staticmethods below read it from ctx
        return super().apply(*args, cls.PRIMITIVE, **kwargs)

    @staticmethod
    def forward(ctx, vertex, center2D, shs, feature, opacity, raster_settings: TriangleRasterizationSettings, primitive="2D"):
        s = raster_settings
        shard = _shard()
        ctx.primitive = primitive
        if primitive == "3D":  # the 3D package makes its inputs contiguous in C++ (R3D/src/extension_interface.cu:82-92) instead of raising
            vertex, shs, feature, opacity = vertex.contiguous(), shs.contiguous(), feature.contiguous(), opacity.contiguous()
        args = (
            s.image_width, s.image_height, s.tanfovx, s.tanfovy, s.viewmatrix.contiguous(), s.projmatrix.contiguous(), s.campos.contiguous(),
            s.sh_degree, s.gamma, s.scale_modifier, float(s.background_depth), s.background.contiguous(), vertex, shs, feature, opacity,
            s.back_culling, s.rich_info, s.debug,
        )
        (num_rendered, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer,
         imageBuffer) = debug_run(_C.rasterize_triangles, *args, debug=s.debug, shard=shard, primitive=primitive)

        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.shard = shard
        ctx.save_for_backward(vertex, shs, feature, opacity, radii, geometryBuffer, binningBuffer, imageBuffer)
        ctx.mark_non_differentiable(radii)
        if s.rich_info:
            ctx.mark_non_differentiable(contrib_sum, contrib_max)
            if shard[1] > 1 and not getattr(out_feature, "_ts2d_assembled", False):
                from . import distributed

                distributed.assemble_forward(out_feature, depth, normal, contrib_sum, contrib_max)
            return out_feature, radii, depth, normal, contrib_sum, contrib_max
        if shard[1] > 1 and not getattr(out_feature, "_ts2d_assembled", False):
            from . import distributed

            distributed.assemble_forward(out_feature)
        return out_feature, radii

    @staticmethod
    def backward(ctx, *grads_out):
        s = ctx.raster_settings
        vertex, shs, feature, opacity, radii, geometryBuffer, binningBuffer, imageBuffer = ctx.saved_tensors
        if s.rich_info:
            grad_out_feature, _, grad_out_depth, grad_out_normal, _, _ = grads_out
            grad_out_depth = grad_out_depth.contiguous()
            grad_out_normal = grad_out_normal.contiguous()
        else:  # the reference raises UnboundLocalError here (__init__.py:116-117,141-142); we support it
            grad_out_feature, _ = grads_out
            grad_out_depth = grad_out_normal = None
        args = (
            s.tanfovx, s.tanfovy, s.viewmatrix.contiguous(), s.projmatrix.contiguous(), s.campos.contiguous(), s.sh_degree, s.gamma,
            s.scale_modifier, float(s.background_depth), s.background.contiguous(), vertex, shs, feature, opacity, ctx.num_rendered, radii,
            geometryBuffer, binningBuffer, imageBuffer, grad_out_feature.contiguous(), grad_out_depth, grad_out_normal, s.rich_info, s.debug,
        )
        grad_vertex, grad_center2D, grad_shs, grad_feature, grad_opacity = debug_run(
            _C.rasterize_triangles_backward, *args, debug=s.debug, shard=ctx.shard, primitive=ctx.primitive)
        # tile-sharded runs: _C.rasterize_triangles_backward already all-reduced the per-triangle accumulators between the
        # composite and the per-triangle stage, so the five gradients are complete and identical on every rank here
        use_shs = feature.dim() <= 1 or (feature.size(0) == 0 and shs.size(0) > 0)
        return (grad_vertex, grad_center2D, grad_shs if use_shs else None, None if use_shs else grad_feature, grad_opacity, None, None)


class _RasterizeTriangles3D(_RasterizeTriangles):
    """Stands in for diff_triangle_rasterization_3D._RasterizeTriangles (R3D/diff_triangle_rasterization_3D/__init__.py:49-164):
    same arguments and return tuples; the per-pixel primitive is the ray / triangle-plane intersection of R3D/src/forward.cu:243-276."""

    PRIMITIVE = "3D"


class TriangleRasterizer(nn.Module):
    _function = _RasterizeTriangles

    def __init__(self, raster_settings: TriangleRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, vertex: torch.Tensor, center2D: torch.Tensor, opacity: torch.Tensor, shs: torch.Tensor = None,
                feature: torch.Tensor = None):
        if (shs is None and feature is None) or (shs is not None and feature is not None):
            raise Exception("Please provide excatly one of either SHs or feature!")
        shs = torch.Tensor([]) if shs is None else shs
        feature = torch.Tensor([]) if feature is None else feature
        return self._function.apply(vertex, center2D, shs, feature, opacity, self.raster_settings)


class TriangleRasterizer3D(TriangleRasterizer):
    """diff_triangle_rasterization_3D.TriangleRasterizer (R3D/diff_triangle_rasterization_3D/__init__.py:167-187)."""

    _function = _RasterizeTriangles3D


from .frontend import TriangleModelRasterizer, TrainingStatistics, bilinear_downsample, gamma_rescale_ratio  # noqa: E402

__all__ += ["TriangleModelRasterizer", "TrainingStatistics", "bilinear_downsample", "gamma_rescale_ratio"]
from .loss import DepthNormalLoss, ImageLoss, depth_normal_loss, image_loss  # noqa: E402

__all__ += ["ImageLoss", "image_loss", "DepthNormalLoss", "depth_normal_loss"]
