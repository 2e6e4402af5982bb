"""CPU restatement of the pixel-wise image loss of the reference trainer (numpy).

TEST INFRASTRUCTURE ONLY -- like everything under oracle/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

Follows src/diff_recon/trainers/trainer_utils.py: GaussianSmoothing2D._gaussian_kernel (:17-30) and .forward (:32-43: F.conv2d with
padding (k-1)//2, groups = channels), SSIM.forward (:56-82), SSIMLoss (:101-103), L1 (:323-324); and the weighting of
VanillaTS_trainer.py:71-75,108.  Pinning: tests/test_loss_oracle.py runs those lines through torch on the CPU (forward in fp32 and
fp64, gradient through autograd) and requires this restatement to reproduce them.  The backward here is written the way the CUDA
kernel computes it (three derivative maps + the same window), so the test also validates that derivation.
"""
from __future__ import annotations

import numpy as np

KERNEL_SIZE, SIGMA = 11, 1.5  # SSIM.__init__ defaults (:46)
C1, C2 = 0.01 ** 2, 0.03 ** 2  # :51-52


def gaussian_kernel_2d(dtype=np.float64) -> np.ndarray:
    """:17-30 -- exp(-((i - mean)^2 + (j - mean)^2) / (2 sigma^2)), normalised by its sum."""
    k = np.arange(KERNEL_SIZE, dtype=np.float64)
    mean = (KERNEL_SIZE - 1) / 2.0
    g = np.exp(-((k[None, :] - mean) ** 2 + (k[:, None] - mean) ** 2) / (2 * SIGMA ** 2.0))
    return (g / g.sum()).astype(dtype)


def smooth(x: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """:32-43 -- depth-wise F.conv2d (a correlation) with zero padding (k-1)//2; x is (planes, H, W)."""
    r = (kernel.shape[0] - 1) // 2
    p, h, w = x.shape
    xp = np.zeros((p, h + 2 * r, w + 2 * r), dtype=x.dtype)
    xp[:, r:r + h, r:r + w] = x
    out = np.zeros_like(x)
    for i in range(kernel.shape[0]):
        for j in range(kernel.shape[1]):
            out += kernel[i, j] * xp[:, i:i + h, j:j + w]
    return out


def ssim_map(img1: np.ndarray, img2: np.ndarray, dtype=np.float64):
    """:56-80 -- returns (ssim_map, intermediates)."""
    x, y = np.asarray(img1, dtype=dtype), np.asarray(img2, dtype=dtype)
    k = gaussian_kernel_2d(dtype)
    mu1, mu2 = smooth(x, k), smooth(y, k)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = smooth(x * x, k) - mu1_sq
    sigma2_sq = smooth(y * y, k) - mu2_sq
    sigma12 = smooth(x * y, k) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return m, dict(mu1=mu1, mu2=mu2, s11=sigma1_sq, s22=sigma2_sq, s12=sigma12)


def image_loss(image: np.ndarray, gt: np.ndarray, w_l1: float, w_ssim: float, dtype=np.float64):
    """w_l1 * L1 (:323-324) + w_ssim * (1 - ssim.mean()) (:81,:103) -> (loss, [L1, 1 - SSIM])."""
    x, y = np.asarray(image, dtype=dtype), np.asarray(gt, dtype=dtype)
    l1 = np.abs(x - y).mean()
    ssim_loss = 1.0 - ssim_map(x, y, dtype)[0].mean()
    return float(w_l1 * l1 + w_ssim * ssim_loss), np.array([l1, ssim_loss])


def image_loss_backward(image: np.ndarray, gt: np.ndarray, w_l1: float, w_ssim: float) -> np.ndarray:
    """d loss / d image (fp64), computed like the CUDA backward: with A = 2 mu1 mu2 + C1, B = 2 s12 + C2, D = mu1^2 + mu2^2 + C1,
    E = s11 + s22 + C2, ssim = A B / (D E):  d1 = d ssim / d mu1 at fixed RAW moments, d2 = d ssim / d E[x^2], d3 = d ssim / d E[xy];
    d loss / d x = -w_ssim / N * (W*d1 + 2 x (W*d2) + y (W*d3)) + w_l1 / N * sign(x - y), W* = the (symmetric) window, zero padded."""
    x, y = np.asarray(image, dtype=np.float64), np.asarray(gt, dtype=np.float64)
    m, t = ssim_map(x, y)
    mu1, mu2, s11, s22, s12 = t["mu1"], t["mu2"], t["s11"], t["s22"], t["s12"]
    A, B = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    D, E = mu1 * mu1 + mu2 * mu2 + C1, s11 + s22 + C2
    g_s11, g_s12 = -m / E, 2 * A / (D * E)
    g_mu1 = 2 * mu2 * B / (D * E) - 2 * mu1 * m / D
    d1, d2, d3 = g_mu1 - 2 * mu1 * g_s11 - mu2 * g_s12, g_s11, g_s12
    k = gaussian_kernel_2d()
    n = x.size
    return -w_ssim / n * (smooth(d1, k) + 2 * x * smooth(d2, k) + y * smooth(d3, k)) + w_l1 / n * np.sign(x - y)
