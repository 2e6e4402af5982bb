"""Fused loss front-end.  depth_normal_loss / DepthNormalLoss: the geometry term (see the function below).
image_loss / ImageLoss: the pixel-wise core of the reference trainer's loss,

    w_L1 * L1(image, gt_image) + w_ssim * ssimLoss(image, gt_image)
    (src/diff_recon/trainers/VanillaTS_trainer.py:74-75,108; trainer_utils.py:323-324 L1, :9-103 GaussianSmoothing2D / SSIM / SSIMLoss)

as one CUDA forward and one CUDA backward kernel (libts2d: ts2d_image_loss_forward / _backward) instead of the ~30 torch kernels
and ten depth-wise 11x11 convolutions the torch composition launches per step.  Same window (11x11 Gaussian, sigma 1.5, normalised,
zero padding), same constants, same mean reductions; fp32 with an IEEE divide.  The reference's own convolutions run through cuDNN
with TF32 allowed (torch's default), so this op is *more* accurate than the reference flow, not less.
No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._C import _ptr, _require_cuda_f32

__all__ = ["image_loss", "ImageLoss", "depth_normal_loss", "DepthNormalLoss"]


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, w_l1, w_ssim):
        lib = _lib.load()
        for name, t in (("image", image), ("gt_image", gt)):
            _require_cuda_f32(name, t)
        if image.shape != gt.shape:
            raise ValueError("Input images must have the same dimensions.")  # trainer_utils.py:84-85
        ctx.in_shape = image.shape  # the gradient goes back in the caller's shape, not in the flattened (planes, H, W) one
        if image.dim() == 2:
            image, gt = image.unsqueeze(0), gt.unsqueeze(0)
        if image.dim() == 4:  # (B, C, H, W): every plane is an independent SSIM plane and both means run over all of them
            image, gt = image.reshape(-1, *image.shape[2:]), gt.reshape(-1, *gt.shape[2:])
        if image.dim() != 3:
            raise ValueError("Input images must have 2, 3, or 4 dimensions.")  # :97-98
        image, gt = image.contiguous(), gt.contiguous()
        ch, h, w = image.shape
        dev = image.device
        sbytes = lib.ts2d_image_loss_scratch_bytes(ch, w, h)
        scratch = torch.empty((sbytes,), device=dev, dtype=torch.uint8)
        loss = torch.empty((1,), device=dev, dtype=torch.float32)
        terms = torch.empty((2,), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.ts2d_image_loss_forward(_ptr(image), _ptr(gt), ch, w, h, float(w_l1), float(w_ssim), _ptr(loss), _ptr(terms),
                                                   _ptr(scratch), sbytes, stream), "ts2d_image_loss_forward")
        ctx.save_for_backward(image, gt, scratch)
        ctx.w = (float(w_l1), float(w_ssim))
        ctx.mark_non_differentiable(terms)
        return loss.reshape(()), terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        lib = _lib.load()
        image, gt, scratch = ctx.saved_tensors
        ch, h, w = image.shape
        dev = image.device
        out = torch.empty_like(image)
        g = g_loss.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.ts2d_image_loss_backward(_ptr(image), _ptr(gt), ch, w, h, ctx.w[0], ctx.w[1], _ptr(g), _ptr(scratch), scratch.numel(),
                                                    _ptr(out), stream), "ts2d_image_loss_backward")
        return out.view(ctx.in_shape), None, None, None


def image_loss(image: torch.Tensor, gt_image: torch.Tensor, w_ssim: float, w_l1: float | None = None, return_terms: bool = False):
    """w_l1 * L1(image, gt_image) + w_ssim * (1 - SSIM(image, gt_image)); w_l1 defaults to 1 - w_ssim (VanillaTS_trainer.py:71).
    Gradient flows to `image` only (the ground truth is data).  `return_terms` adds the detached tensor [L1, 1 - SSIM]."""
    w_l1 = 1.0 - float(w_ssim) if w_l1 is None else float(w_l1)
    loss, terms = _ImageLoss.apply(image, gt_image, w_l1, float(w_ssim))
    return (loss, terms) if return_terms else loss


class ImageLoss(torch.nn.Module):
    """Module form, for trainers that hold their losses as attributes (VanillaTS_trainer.py:26-31)."""

    def __init__(self, w_ssim: float, w_l1: float | None = None):
        super().__init__()
        self.w_ssim, self.w_l1 = float(w_ssim), w_l1

    def forward(self, image: torch.Tensor, gt_image: torch.Tensor) -> torch.Tensor:
        return image_loss(image, gt_image, self.w_ssim, self.w_l1)


class _DepthNormalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, normal, tan_fovx, tan_fovy, half, quantile, depth_grad, normal_grad):
        lib = _lib.load()
        _require_cuda_f32("depth", depth)
        _require_cuda_f32("normal", normal)
        if depth.dim() != 2 or normal.dim() != 3 or normal.shape[0] != 3 or tuple(normal.shape[1:]) != tuple(depth.shape):
            raise ValueError("depth must be (H, W) and normal (3, H, W)")
        depth, normal = depth.contiguous(), normal.contiguous()
        h, w = depth.shape
        dev = depth.device
        sbytes = lib.ts2d_depth_normal_loss_scratch_bytes(w, h, int(half))
        if sbytes == 0:
            raise RuntimeError(_lib.error_string(-11))
        scratch = torch.empty((sbytes,), device=dev, dtype=torch.uint8)
        loss = torch.empty((1,), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.ts2d_depth_normal_loss_forward(_ptr(depth), _ptr(normal), w, h, float(tan_fovx), float(tan_fovy), int(half), float(quantile),
                                                          _ptr(loss), _ptr(scratch), sbytes, stream), "ts2d_depth_normal_loss_forward")
        ctx.save_for_backward(scratch)
        ctx.meta = (h, w, float(tan_fovx), float(tan_fovy), int(half), bool(depth_grad), bool(normal_grad))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        (scratch,) = ctx.saved_tensors
        h, w, tfx, tfy, half, depth_grad, normal_grad = ctx.meta
        dev = scratch.device
        want_d, want_n = depth_grad and ctx.needs_input_grad[0], normal_grad and ctx.needs_input_grad[1]
        g_depth = torch.empty((h, w), device=dev, dtype=torch.float32) if want_d else None
        g_normal = torch.empty((3, h, w), device=dev, dtype=torch.float32) if want_n else None
        g = g_loss.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.ts2d_depth_normal_loss_backward(w, h, tfx, tfy, half, _ptr(g), _ptr(scratch), scratch.numel(), _ptr(g_depth), _ptr(g_normal),
                                                           stream), "ts2d_depth_normal_loss_backward")
        return g_depth, g_normal, None, None, None, None, None, None


def depth_normal_loss(depth: torch.Tensor, normal: torch.Tensor, tan_fovx: float, tan_fovy: float, scale_factor: float | None = None,
                      depth_grad: bool = True, normal_grad: bool = True, depth_grad_filter_quantile: float = 0.9) -> torch.Tensor:
    """The geometry term of the reference trainer, DepthNormalLoss(...)(depth, normal, tan_fovx, tan_fovy)
    (src/diff_recon/trainers/trainer_utils.py:203-257; VanillaTS_trainer.py:84), as fused kernels (libts2d: ts2d_depth_normal_loss_*):
    mean((1 - <normalize(normal), normal_from_depth(depth)>) * mask).  `scale_factor` None / 1 or 0.5 (the reference's shipped value)."""
    if scale_factor is None or scale_factor == 1:
        half = 0
    elif scale_factor == 0.5:
        half = 1
    else:
        raise ValueError("depth_normal_loss supports scale_factor None, 1 or 0.5")
    return _DepthNormalLoss.apply(depth, normal, tan_fovx, tan_fovy, half, depth_grad_filter_quantile, depth_grad, normal_grad)


class DepthNormalLoss(torch.nn.Module):
    """Module form with the reference's constructor (trainer_utils.py:204-210) and call signature (:249)."""

    def __init__(self, depth_grad: bool = True, normal_grad: bool = True, scale_factor: float | None = None, depth_grad_filter_quantile: float = 0.9):
        super().__init__()
        self.depth_grad, self.normal_grad = depth_grad, normal_grad
        self.scale_factor, self.depth_grad_filter_quantile = scale_factor, depth_grad_filter_quantile

    def forward(self, depth: torch.Tensor, normal: torch.Tensor, tan_fovx: float, tan_fovy: float) -> torch.Tensor:
        return depth_normal_loss(depth, normal, tan_fovx, tan_fovy, self.scale_factor, self.depth_grad, self.normal_grad,
                                 self.depth_grad_filter_quantile)
