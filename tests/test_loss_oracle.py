"""CPU: pins oracle/loss.py (restatement of the trainer's pixel-wise image loss) against torch running
the reference's own lines (src/diff_recon/trainers/trainer_utils.py:9-103,323-324), copied below."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import loss as lo  # noqa: E402


def _reference_kernel(kernel_size=11, sigma=1.5):
    # trainer_utils.py:17-30, verbatim
    x_grid = torch.arange(kernel_size).unsqueeze(0).repeat(kernel_size, 1)
    xy_grid = torch.stack([x_grid, x_grid.T], dim=-1).float()
    mean = (kernel_size - 1) / 2.0
    variance = sigma**2.0
    kernel = torch.exp(-(xy_grid - mean).pow(2).sum(dim=-1) / (2 * variance))
    kernel = kernel / kernel.sum()
    return kernel.unsqueeze(0).unsqueeze(0).float()


def reference_loss(image, gt, w_l1, w_ssim):
    """L1 (:323-324) and SSIMLoss (:56-82, :101-103) on (C, H, W) tensors, in the dtype of the inputs."""
    img1, img2 = image.unsqueeze(0), gt.unsqueeze(0)  # normalize_shape :88-90
    channels = img1.shape[1]
    kernel = _reference_kernel().to(img1.dtype).repeat(channels, 1, 1, 1).to(img1.device)
    window = lambda x: F.conv2d(x, kernel, padding=5, groups=channels)  # :32-43
    mu1, mu2 = window(img1), window(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = window(img1 * img1) - mu1_sq
    sigma2_sq = window(img2 * img2) - mu2_sq
    sigma12 = window(img1 * img2) - mu1_mu2
    C1, C2 = 0.01**2, 0.03**2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    ssim_loss = 1 - ssim_map.mean()
    l1 = torch.abs((image - gt)).mean()
    return w_l1 * l1 + w_ssim * ssim_loss, l1, ssim_loss


@pytest.fixture(scope="module")
def images():
    g = torch.Generator().manual_seed(3)
    h, w = 37, 53  # not multiples of the 16-pixel kernel tile, smaller than 2 tiles + halo in one direction
    gt = torch.rand(3, h, w, generator=g, dtype=torch.float64)
    img = (gt + 0.15 * torch.randn(3, h, w, generator=g, dtype=torch.float64)).clamp(0, 1)
    img[:, 5:9, 7:12] = gt[:, 5:9, 7:12]  # a region of exact agreement: sign(0) = 0 in the L1 gradient
    return img, gt


def test_window_is_the_reference_kernel_and_separable():
    k = lo.gaussian_kernel_2d(np.float32)
    assert np.allclose(k, _reference_kernel()[0, 0].numpy(), rtol=2e-6, atol=0)  # the reference builds it in fp32 (exp and sum)
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    assert np.allclose(lo.gaussian_kernel_2d(), np.outer(g, g) / g.sum() ** 2, rtol=1e-14)  # what the CUDA kernels rely on


@pytest.mark.parametrize("w_ssim", [0.2, 0.0, 1.0])
def test_forward_matches_reference_lines(images, w_ssim):
    img, gt = images
    ref, l1, sl = reference_loss(img, gt, 1 - w_ssim, w_ssim)
    mine, terms = lo.image_loss(img.numpy(), gt.numpy(), 1 - w_ssim, w_ssim)
    # the reference builds its window in fp32 (relative error ~1e-7 per weight): that is the only difference to the fp64 restatement
    assert abs(mine - float(ref)) <= 2e-8 and abs(terms[0] - float(l1)) <= 1e-13 and abs(terms[1] - float(sl)) <= 5e-8
    ref32 = reference_loss(img.float(), gt.float(), 1 - w_ssim, w_ssim)[0]
    mine32 = lo.image_loss(img.float().numpy(), gt.float().numpy(), 1 - w_ssim, w_ssim, dtype=np.float32)[0]
    assert abs(mine32 - float(ref32)) <= 2e-6 * max(1.0, abs(float(ref32)))


@pytest.mark.parametrize("w_ssim", [0.2, 1.0])
def test_backward_matches_autograd(images, w_ssim):
    img, gt = images
    x = img.clone().requires_grad_(True)
    reference_loss(x, gt, 1 - w_ssim, w_ssim)[0].backward()
    mine = lo.image_loss_backward(img.numpy(), gt.numpy(), 1 - w_ssim, w_ssim)
    scale = np.abs(x.grad.numpy()).max()
    assert np.abs(mine - x.grad.numpy()).max() <= 2e-6 * scale  # fp32-built window of the reference vs the fp64 one here
