"""GPU parity of the fused image loss (first step) against the CPU oracle (fp64) and against torch running
the reference's own lines (trainer_utils.py:9-103,323-324) on the same device with TF32 off.  Bars (fp32 kernel; the E[x^2] - mu^2
cancellation of the SSIM definition limits fp32 -- torch's fp32 composition has the same error, checked below): loss 1e-5 relative,
gradient max error 1e-3 and mean error 5e-5 of the largest gradient entry."""
import numpy as np
import pytest
import torch

import harness  # noqa: F401  (sys.path)
from test_loss_oracle import reference_loss

pytestmark = pytest.mark.gpu

LOSS_REL, GRAD_MAX, GRAD_MEAN = 1e-5, 1e-3, 5e-5


def _pair(c, h, w, seed, smooth=False):
    g = torch.Generator().manual_seed(seed)
    if smooth:
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
        gt = torch.stack([0.5 + 0.4 * torch.sin(xx / 30 + k) * torch.cos(yy / 25) for k in range(c)])
        img = (gt + 0.02 * torch.randn(c, h, w, generator=g, dtype=torch.float64)).clamp(0, 1)
    else:
        gt = torch.rand(c, h, w, generator=g, dtype=torch.float64)
        img = (gt + 0.15 * torch.randn(c, h, w, generator=g, dtype=torch.float64)).clamp(0, 1)
    if h > 8 and w > 12:
        img[:, 5:8, 7:12] = gt[:, 5:8, 7:12]  # exact agreement: sign(0) = 0
    return img, gt


def _ours(img, gt, w_ssim, dev, scale=1.0):
    from triangle_splatting_b200 import image_loss

    x = img.float().to(dev).requires_grad_(True)
    loss, terms = image_loss(x, gt.float().to(dev), w_ssim, return_terms=True)
    (scale * loss).backward()
    return float(loss.detach()), terms.cpu().numpy(), x.grad.cpu().numpy()


@pytest.mark.parametrize("c,h,w", [(3, 37, 53), (3, 16, 16), (1, 4, 7), (2, 40, 17), (3, 1, 1), (3, 64, 96)])
@pytest.mark.parametrize("w_ssim", [0.2, 1.0])
def test_image_loss_vs_oracle(c, h, w, w_ssim, cuda_device):
    from oracle import loss as lo

    img, gt = _pair(c, h, w, seed=h * 100 + w)
    img, gt = img.float().double(), gt.float().double()  # the values the kernel sees
    loss, terms, grad = _ours(img, gt, w_ssim, cuda_device)
    ref, rterms = lo.image_loss(img.numpy(), gt.numpy(), 1 - w_ssim, w_ssim)
    rgrad = lo.image_loss_backward(img.numpy(), gt.numpy(), 1 - w_ssim, w_ssim)
    assert abs(loss - ref) <= LOSS_REL * abs(ref)
    assert np.allclose(terms, rterms, rtol=LOSS_REL, atol=1e-7)
    scale = np.abs(rgrad).max()
    assert np.abs(grad - rgrad).max() <= GRAD_MAX * scale and np.abs(grad - rgrad).mean() <= GRAD_MEAN * scale


@pytest.mark.parametrize("smooth", [False, True])
def test_image_loss_vs_torch_reference_lines(smooth, cuda_device):
    """A 480x270 frame: ours vs the reference's lines in torch fp64 (truth) and fp32 (what the trainer runs, TF32 off here)."""
    dev = cuda_device
    img, gt = _pair(3, 270, 480, seed=1, smooth=smooth)
    img, gt = img.float().double(), gt.float().double()
    w_ssim = 0.2
    x64 = img.to(dev).requires_grad_(True)
    t64 = reference_loss(x64, gt.to(dev), 1 - w_ssim, w_ssim)[0]
    t64.backward()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x32 = img.float().to(dev).requires_grad_(True)
        t32 = reference_loss(x32, gt.float().to(dev), 1 - w_ssim, w_ssim)[0]
        t32.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    loss, _, grad = _ours(img, gt, w_ssim, dev)
    truth, g64, g32 = float(t64), x64.grad.cpu().numpy(), x32.grad.cpu().numpy().astype(np.float64)
    scale = np.abs(g64).max()
    assert abs(loss - truth) <= LOSS_REL * abs(truth)
    err_ours, err_t32 = np.abs(grad - g64), np.abs(g32 - g64)
    assert err_ours.max() <= GRAD_MAX * scale and err_ours.mean() <= GRAD_MEAN * scale
    assert err_ours.mean() <= 3 * err_t32.mean() + 1e-7 * scale, "less accurate than torch's own fp32 composition"


def test_image_loss_upstream_gradient_batches_and_errors(cuda_device):
    from triangle_splatting_b200 import ImageLoss, image_loss

    dev = cuda_device
    img, gt = _pair(3, 33, 47, seed=9)
    l1, _, g1 = _ours(img, gt, 0.2, dev)
    l2, _, g2 = _ours(img, gt, 0.2, dev, scale=2.5)
    assert abs(l1 - l2) <= 1e-6 * abs(l1) and np.allclose(g2, 2.5 * g1, rtol=1e-5, atol=1e-12)
    # (B, C, H, W): every plane is its own SSIM plane, both means run over all of them (normalize_shape, trainer_utils.py:86-87)
    b_img = torch.stack([img, img.flip(-1)]).float().to(dev)
    b_gt = torch.stack([gt, gt.flip(-1)]).float().to(dev)
    lb = float(image_loss(b_img, b_gt, 0.2))
    assert abs(lb - l1) <= 1e-5 * abs(l1)  # flipping both images leaves the loss unchanged
    assert abs(float(ImageLoss(0.2)(img.float().to(dev), gt.float().to(dev))) - l1) <= 1e-6 * abs(l1)
    # 2-D input
    l2d = float(image_loss(img[0].float().to(dev), gt[0].float().to(dev), 0.2))
    from oracle import loss as lo

    assert abs(l2d - lo.image_loss(img[:1].float().double().numpy(), gt[:1].float().double().numpy(), 0.8, 0.2)[0]) <= 1e-5 * l2d
    with pytest.raises(RuntimeError):
        image_loss(img.float(), gt.float().to(dev), 0.2)  # no CPU path
    with pytest.raises(ValueError):
        image_loss(img.float().to(dev), gt[:, :-1].float().to(dev), 0.2)


@pytest.mark.parametrize("shape", ["2d", "4d"])
def test_image_loss_backward_keeps_the_callers_shape(shape, cuda_device):
    """(H, W) and (B, C, H, W) inputs: the gradient comes back in the input's own shape and equals the (planes, H, W) result."""
    from triangle_splatting_b200 import image_loss

    dev = cuda_device
    img, gt = _pair(3, 33, 47, seed=21)
    if shape == "2d":
        x_in, y_in = img[0].float().to(dev), gt[0].float().to(dev)
        x3, y3 = img[:1].float().to(dev), gt[:1].float().to(dev)
    else:
        x_in = torch.stack([img, img.flip(-1)]).float().to(dev)
        y_in = torch.stack([gt, gt.flip(-1)]).float().to(dev)
        x3, y3 = x_in.reshape(-1, 33, 47).clone(), y_in.reshape(-1, 33, 47).clone()
    x_in.requires_grad_(True)
    x3.requires_grad_(True)
    (3.0 * image_loss(x_in, y_in, 0.2)).backward()
    (3.0 * image_loss(x3, y3, 0.2)).backward()
    assert x_in.grad.shape == x_in.shape
    assert torch.equal(x_in.grad.reshape(x3.shape), x3.grad)
    # and against torch running the reference's lines on the flattened planes
    x64 = x3.detach().double().requires_grad_(True)
    (3.0 * reference_loss(x64, y3.double(), 0.8, 0.2)[0]).backward()
    g64 = x64.grad.cpu().numpy()
    assert np.abs(x_in.grad.reshape(x3.shape).cpu().numpy() - g64).max() <= GRAD_MAX * np.abs(g64).max()
