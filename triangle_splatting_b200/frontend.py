"""Parameter-space front-end of the rasterizer.

``diff_recon``'s model wraps every rasterizer call in a Python preamble and epilogue
(src/diff_recon/models/VanillaTS_model.py):

    :80   shs     = torch.cat((_f_dc, _f_rest), dim=1)                 576 MB of traffic per frame at 1.5 M triangles, SH 3
    :84   opacity = torch.sigmoid(_opacity)
    :615-618 + :431-447  vertex_rescale = (vertex - mean) * ratio + mean     (gamma_rescale)
    :620-621  opacity_ste = ((opacity > thr).float() - opacity).detach() + opacity
    :623  bg_depth = (camera_center - vertex).norm(dim=-1).max()         -> float(bg_depth): a device->host sync per frame
    :625-630  render at render_up_scale x the camera resolution
    :647-655  F.interpolate(..., mode="bilinear") of render / depth / normal back to (h, w); radii // render_up_scale
    :347-363  _training_statistic: six masked scatter updates per step

``TriangleModelRasterizer`` takes the raw parameters instead and does all of that inside the per-triangle kernels
(K1 / K9 of libts2d) and two small resize kernels -- same operation order as the torch kernels above, so the values that
reach the reference-order geometry code are the ones the reference would have materialised.  The reference-shaped
``TriangleRasterizer`` is untouched: an unchanged trainer keeps working through it; this module is the additional entry
point a maintainer switches ``VanillaTSModel.forward`` to.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _C
from . import TriangleRasterizationSettings, _shard, debug_run

__all__ = ["TriangleModelRasterizer", "TrainingStatistics", "gamma_rescale_ratio", "bilinear_downsample"]


def gamma_rescale_ratio(gamma: float) -> float:
    """VanillaTS_model.py:616-617: beta = 1/gamma; ratio = 1 / sqrt(2**beta * beta * Gamma(beta))."""
    import math

    beta = 1.0 / float(gamma)
    return 1.0 / math.sqrt(2.0 ** beta * beta * math.gamma(beta))


class TrainingStatistics:
    """The six per-triangle accumulators of VanillaTS_model.py:196-201, updated in place by the backward pass
    (:347-363) for the triangles with radii > 0."""

    FIELDS = _C.STAT_FIELDS

    def __init__(self, num_points: int, device):
        for k in self.FIELDS:
            setattr(self, k, torch.zeros((num_points,), device=device, dtype=torch.float32))

    def as_dict(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, k) for k in self.FIELDS}


class _Downsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return _C.downsample(x.contiguous(), s)

    @staticmethod
    def backward(ctx, g):
        return _C.downsample(g.contiguous(), ctx.s, backward=True), None


def bilinear_downsample(x: torch.Tensor, s: int) -> torch.Tensor:
    """F.interpolate(x[None], size=(H // s, W // s), mode="bilinear")[0] for a planar (planes, H, W) CUDA tensor."""
    return x if s == 1 else _Downsample.apply(x, int(s))


class _RasterizeModel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertex, center2D, f_dc, f_rest, opacity_logit, raster_settings: TriangleRasterizationSettings, opts: dict):
        s = raster_settings
        shard = _shard()
        model = _C.ModelInputs(f_dc.contiguous(), None if f_rest is None else f_rest.contiguous(), opacity_logit.contiguous(),
                               opts["ste_threshold"], opts["rescale_ratio"], opts["bg_depth_from_vertices"])
        vertex = vertex.contiguous()
        args = (
            s.image_width, s.image_height, s.tanfovx, s.tanfovy, s.viewmatrix.contiguous(), s.projmatrix.contiguous(), s.campos.contiguous(),
            s.sh_degree, s.gamma, s.scale_modifier, 0.0 if opts["bg_depth_from_vertices"] else float(s.background_depth),
            s.background.contiguous(), vertex, None, None, None, s.back_culling, s.rich_info, s.debug,
        )
        (num_rendered, out_feature, radii, depth, normal, contrib_sum, contrib_max, geometryBuffer, binningBuffer,
         imageBuffer) = debug_run(_C.rasterize_triangles, *args, debug=s.debug, shard=shard, primitive=opts["primitive"], model=model)
        ctx.raster_settings, ctx.num_rendered, ctx.shard, ctx.opts = s, num_rendered, shard, opts
        ctx.has_rest = f_rest is not None
        saved = [vertex, model.f_dc, model.opacity_logit, radii, geometryBuffer, binningBuffer, imageBuffer, contrib_sum, contrib_max]
        if ctx.has_rest:
            saved.append(model.f_rest)
        ctx.save_for_backward(*saved)
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
        s, opts = ctx.raster_settings, ctx.opts
        vertex, f_dc, logit, radii, geometryBuffer, binningBuffer, imageBuffer, contrib_sum, contrib_max = ctx.saved_tensors[:9]
        f_rest = ctx.saved_tensors[9] if ctx.has_rest else None
        if s.rich_info:
            g_feature, _, g_depth, g_normal, _, _ = grads_out
            g_depth, g_normal = g_depth.contiguous(), g_normal.contiguous()
        else:
            g_feature, _ = grads_out
            g_depth = g_normal = None
        model = _C.ModelInputs(f_dc, f_rest, logit, opts["ste_threshold"], opts["rescale_ratio"], opts["bg_depth_from_vertices"])
        stats = opts["statistics"]
        args = (
            s.tanfovx, s.tanfovy, s.viewmatrix.contiguous(), s.projmatrix.contiguous(), s.campos.contiguous(), s.sh_degree, s.gamma,
            s.scale_modifier, 0.0 if opts["bg_depth_from_vertices"] else float(s.background_depth), s.background.contiguous(), vertex, None,
            None, None, ctx.num_rendered, radii, geometryBuffer, binningBuffer, imageBuffer, g_feature.contiguous(), g_depth, g_normal,
            s.rich_info, s.debug,
        )
        gv, gc, gdc, grest, gop = debug_run(
            _C.rasterize_triangles_backward, *args, debug=s.debug, shard=ctx.shard, primitive=opts["primitive"], model=model,
            stats=stats.as_dict() if isinstance(stats, TrainingStatistics) else stats,
            fwd_contrib=(contrib_sum, contrib_max) if s.rich_info else None, radii_div=opts["radii_div"])
        return gv, gc, gdc, (grest if ctx.has_rest else None), gop.view_as(logit), None, None


class TriangleModelRasterizer(nn.Module):
    """Rasterizer fed with the model's raw parameters.

    raster_settings : the same NamedTuple the reference-shaped rasterizer takes, at the CAMERA resolution (w, h);
                      ``background_depth`` is ignored when ``bg_depth_from_vertices`` (the model's own choice, :623).
    primitive       : "2D" | "3D" (``rasterizer_type`` of the reference's TriangleRenderer).
    ste_threshold   : config.model.ste_threshold (None = off).
    rescale_ratio   : gamma_rescale_ratio(gamma) when config.model.gamma_rescale, else 1.
    render_up_scale : config.model.render_up_scale; render / depth / normal come back at (h, w), radii // scale.
    statistics      : TrainingStatistics (or a dict with any subset of its six tensors) updated in place by backward();
                      pass None outside the statistic window (VanillaTS_model.py:349-350).

    forward(vertex, center2D, opacity, f_dc, f_rest) -> the reference's tuple: (render, radii) or
    (render, radii, depth, normal, contrib_sum, contrib_max).
    """

    def __init__(self, raster_settings: TriangleRasterizationSettings, *, primitive: str = "2D", ste_threshold: Optional[float] = None,
                 rescale_ratio: float = 1.0, render_up_scale: int = 1, bg_depth_from_vertices: bool = True, statistics=None):
        super().__init__()
        if primitive not in ("2D", "3D"):
            raise ValueError(f"Unknown rasterizer type: {primitive}. Use '2D' or '3D'.")
        if int(render_up_scale) != render_up_scale or render_up_scale < 1:
            raise ValueError("render_up_scale must be a positive integer")
        self.raster_settings = raster_settings
        self.primitive, self.ste_threshold, self.rescale_ratio = primitive, ste_threshold, float(rescale_ratio)
        self.render_up_scale, self.bg_depth_from_vertices, self.statistics = int(render_up_scale), bool(bg_depth_from_vertices), statistics

    def forward(self, vertex: torch.Tensor, center2D: torch.Tensor, opacity: torch.Tensor, f_dc: torch.Tensor,
                f_rest: Optional[torch.Tensor] = None):
        s, k = self.raster_settings, self.render_up_scale
        if k > 1:
            s = s._replace(image_width=int(s.image_width) * k, image_height=int(s.image_height) * k)
        opts = dict(primitive=self.primitive, ste_threshold=self.ste_threshold, rescale_ratio=self.rescale_ratio,
                    bg_depth_from_vertices=self.bg_depth_from_vertices, statistics=self.statistics, radii_div=k)
        if f_rest is not None and f_rest.numel() == 0:
            f_rest = None
        out = _RasterizeModel.apply(vertex, center2D, f_dc, f_rest, opacity, s, opts)
        if k == 1:
            return out
        if s.rich_info:
            render, radii, depth, normal, csum, cmax = out
            return (bilinear_downsample(render, k), radii // k, bilinear_downsample(depth.unsqueeze(0), k).squeeze(0),
                    bilinear_downsample(normal, k), csum, cmax)
        render, radii = out
        return bilinear_downsample(render, k), radii // k
