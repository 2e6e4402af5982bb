"""GPU parity of the fused depth-normal consistency loss (the geometry term of the trainer) against
 (a) the outputs of the reference's own DepthNormalLoss committed as fixtures (tests/golden/depth_normal_*.npz, make_loss_golden.py),
 (b) the CPU oracle oracle/loss.py: depth_normal_loss (fp64) on seeded inputs incl. odd and tiny frames,
 (c) the reference's lines restated with torch ops on the same device at 1080p (TF32 off), incl. torch.quantile's own threshold.
Bars (fp32 kernels): loss 1e-5 relative; gradients: max error 2e-4 of the largest entry, mean error 1e-5 of it.  A pixel whose
gradient norm sits within rounding of the quantile threshold may fall on the other side of the mask: it moves the loss by 1 / N
and one pixel's gradient by its own size, so the gradient bar is stated with a count of such pixels (<= 2 per frame)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import harness  # noqa: F401  (sys.path)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_loss_golden as mk  # noqa: E402

LOSS_REL, GRAD_MAX, GRAD_MEAN, MAX_FLIPS = 1e-5, 2e-4, 1e-5, 2


def _ours(depth, normal, tfx, tfy, sf, dev, scale=1.0, **kw):
    from triangle_splatting_b200 import depth_normal_loss

    d = torch.as_tensor(depth, dtype=torch.float32).to(dev).requires_grad_(True)
    n = torch.as_tensor(normal, dtype=torch.float32).to(dev).requires_grad_(True)
    loss = depth_normal_loss(d, n, tfx, tfy, sf, **kw)
    (scale * loss).backward()
    g = lambda t: None if t.grad is None else t.grad.cpu().numpy()
    return float(loss.detach()), g(d), g(n)


def _check_grad(ours, ref, what):
    """max / mean error against `ref`, after setting aside at most MAX_FLIPS pixels (all their channels) that crossed the mask."""
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max()
    err = np.abs(ours - ref)
    pix = err.reshape(-1, *err.shape[-2:]).max(axis=0)  # worst channel per pixel
    bad = np.argwhere(pix > GRAD_MAX * scale)
    # a flipped pixel also moves its 3x3 / 4x4 stencil neighbours in dL/ddepth: allow the flips' neighbourhoods
    assert len(bad) <= MAX_FLIPS * 40, f"{what}: {len(bad)} pixels above {GRAD_MAX} of the largest entry (max {pix.max() / scale:.2e})"
    keep = pix <= GRAD_MAX * scale
    assert (err.reshape(-1, *err.shape[-2:])[:, keep]).mean() <= GRAD_MEAN * scale, what


@pytest.mark.parametrize("name", ["half_38x50", "half_odd_37x53", "full_21x25", "half_96x128"])
def test_vs_reference_fixtures(name, cuda_device):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"depth_normal_{name}.npz"))
    sf = None if float(z["scale_factor"]) < 0 else float(z["scale_factor"])
    loss, gd, gn = _ours(z["depth"], z["normal"], float(z["tan_fovx"]), float(z["tan_fovy"]), sf, cuda_device)
    assert abs(loss - float(z["loss"])) <= LOSS_REL * abs(float(z["loss"])) + 1.0 / z["depth"].size
    _check_grad(gd, z["g_depth"], f"{name}: dL/ddepth")
    _check_grad(gn, z["g_normal"], f"{name}: dL/dnormal")


@pytest.mark.parametrize("h,w,sf", [(40, 64, 0.5), (41, 63, 0.5), (2, 2, 0.5), (3, 5, 0.5), (1, 1, None), (17, 9, 1), (128, 200, 0.5), (135, 240, None)])
def test_vs_oracle(h, w, sf, cuda_device):
    from oracle import loss as lo

    depth, normal = mk.scene(h, w, seed=31 * h + w)
    loss, gd, gn = _ours(depth, normal, 0.7, 0.5, sf, cuda_device, scale=3.0)
    ref, rgd, rgn = lo.depth_normal_loss(depth.numpy(), normal.numpy(), 0.7, 0.5, sf, with_grad=True)
    assert abs(loss - ref) <= LOSS_REL * abs(ref) + 1.0 / (h * w)
    _check_grad(gd, 3.0 * rgd, f"{h}x{w} sf={sf}: dL/ddepth")
    _check_grad(gn, 3.0 * rgn, f"{h}x{w} sf={sf}: dL/dnormal")


def reference_lines(depth, normal, tan_fovx, tan_fovy, scale_factor, q=0.9):
    """trainer_utils.py:159-185 (ScharrFilter) and :212-255 (DepthNormalLoss) with torch ops, in the dtype / on the device of the inputs."""
    kx = torch.tensor([[-3, 0, 3], [-10, 0, 10], [-3, 0, 3]], dtype=depth.dtype, device=depth.device).view(1, 1, 3, 3) / 32
    ky = torch.tensor([[-3, -10, -3], [0, 0, 0], [3, 10, 3]], dtype=depth.dtype, device=depth.device).view(1, 1, 3, 3) / 32
    W0, H0 = depth.shape[-1], depth.shape[-2]
    d = depth.unsqueeze(0).unsqueeze(0)
    if scale_factor is not None and scale_factor != 1:
        d = F.interpolate(d, scale_factor=scale_factor, mode="bilinear", align_corners=False)
    grad = torch.cat((F.conv2d(d, kx, padding=1), F.conv2d(d, ky, padding=1)), dim=1).squeeze(0)
    Dx, Dy = torch.unbind(grad / d.squeeze(0), 0)
    W, H = d.shape[-1], d.shape[-2]
    x, y = torch.meshgrid(torch.arange(W, dtype=torch.float32, device=depth.device), torch.arange(H, dtype=torch.float32, device=depth.device), indexing="xy")
    nrm = torch.stack([W * Dx / (2 * tan_fovx), H * Dy / (2 * tan_fovy), -(1 + (x - W / 2 + 0.5) * Dx + (y - H / 2 + 0.5) * Dy)], dim=0)
    gnorm = grad.norm(dim=0, keepdim=True)
    if W0 != W or H0 != H:
        nrm = F.interpolate(nrm.unsqueeze(0), size=(H0, W0), mode="bilinear", align_corners=False).squeeze(0)
        gnorm = F.interpolate(gnorm.unsqueeze(0), size=(H0, W0), mode="bilinear", align_corners=False).squeeze(0)
    nrm = nrm / nrm.norm(dim=0, keepdim=True)
    mask = (gnorm < torch.quantile(gnorm, q)).float().squeeze(0)
    normal = F.normalize(normal, p=2, dim=0, eps=1e-8)
    return ((1 - (normal * nrm).sum(dim=0)) * mask).mean()


@pytest.mark.parametrize("sf", [0.5, None])
def test_full_frame_vs_torch_lines_and_determinism(sf, cuda_device):
    """1920x1080: ours vs the reference's lines in torch fp32 on the same device (TF32 off); two runs of ours bit-identical."""
    dev = cuda_device
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        depth, normal = mk.scene(1080, 1920, seed=9)
        d = depth.to(dev).requires_grad_(True)
        n = normal.to(dev).requires_grad_(True)
        ref = reference_lines(d, n, 0.6, 0.34, sf)
        ref.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    loss, gd, gn = _ours(depth, normal, 0.6, 0.34, sf, dev)
    loss2, gd2, gn2 = _ours(depth, normal, 0.6, 0.34, sf, dev)
    assert loss == loss2 and np.array_equal(gd, gd2) and np.array_equal(gn, gn2), "two runs differ"
    assert abs(loss - float(ref)) <= LOSS_REL * abs(float(ref))
    _check_grad(gd, d.grad.cpu().numpy(), "1080p dL/ddepth")
    _check_grad(gn, n.grad.cpu().numpy(), "1080p dL/dnormal")


def test_detached_inputs_and_module_form(cuda_device):
    from triangle_splatting_b200 import DepthNormalLoss

    depth, normal = mk.scene(48, 80, seed=2)
    loss, gd, gn = _ours(depth, normal, 0.6, 0.4, 0.5, cuda_device)
    l2, gd2, gn2 = _ours(depth, normal, 0.6, 0.4, 0.5, cuda_device, depth_grad=False)
    assert l2 == loss and gd2 is None and np.array_equal(gn2, gn)
    l3, gd3, gn3 = _ours(depth, normal, 0.6, 0.4, 0.5, cuda_device, normal_grad=False)
    assert l3 == loss and gn3 is None and np.array_equal(gd3, gd)
    d = depth.to(cuda_device).requires_grad_(True)
    n = normal.to(cuda_device).requires_grad_(True)
    m = DepthNormalLoss(scale_factor=0.5)(d, n, 0.6, 0.4)
    assert float(m.detach()) == loss


def test_argument_errors(cuda_device):
    from triangle_splatting_b200 import depth_normal_loss

    d = torch.ones(8, 8, device=cuda_device)
    n = torch.ones(3, 8, 8, device=cuda_device)
    with pytest.raises(ValueError):
        depth_normal_loss(d, n, 0.5, 0.5, scale_factor=0.25)
    with pytest.raises(ValueError):
        depth_normal_loss(d, n[:2], 0.5, 0.5)
    with pytest.raises(RuntimeError):
        depth_normal_loss(d.cpu(), n.cpu(), 0.5, 0.5)
    with pytest.raises(RuntimeError):
        depth_normal_loss(torch.ones(1, 1, device=cuda_device), torch.ones(3, 1, 1, device=cuda_device), 0.5, 0.5, scale_factor=0.5)
