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


# ------------------------------------------------------------------------------------------------------------------------------
# Depth-normal consistency loss (the trainer's geometry term): trainer_utils.py:159-185 (ScharrFilter), :203-257 (DepthNormalLoss).
# Restated with explicit resampling matrices so that the backward is the transpose of the very same operators; the backward is
# written the way the CUDA kernels compute it (tests/test_loss_oracle.py checks both directions against torch autograd running the
# reference's own class, and against the committed fixtures tests/golden/depth_normal_*.npz generated from it).

SCHARR_X = np.array([[-3, 0, 3], [-10, 0, 10], [-3, 0, 3]], dtype=np.float64) / 32  # :162
SCHARR_Y = np.array([[-3, -10, -3], [0, 0, 0], [3, 10, 3]], dtype=np.float64) / 32  # :163


def correlate3(x: np.ndarray, k: np.ndarray, transpose: bool = False) -> np.ndarray:
    """F.conv2d(x, k, padding=1) of one plane (:178-179), or its adjoint."""
    h, w = x.shape
    xp = np.zeros((h + 2, w + 2), dtype=x.dtype)
    xp[1:1 + h, 1:1 + w] = x
    out = np.zeros_like(x)
    for i in range(3):
        for j in range(3):
            out += (k[2 - i, 2 - j] if transpose else k[i, j]) * xp[i:i + h, j:j + w]
    return out


def down_matrix(n_in: int) -> np.ndarray:
    """F.interpolate(scale_factor=0.5, mode="bilinear", align_corners=False) along one axis (:217): floor(n/2) outputs, output d
    samples the source at 2 (d + 0.5) - 0.5 = 2d + 0.5, i.e. the mean of elements 2d and 2d + 1."""
    n_out = n_in // 2
    m = np.zeros((n_out, n_in))
    for d in range(n_out):
        m[d, 2 * d] = m[d, 2 * d + 1] = 0.5
    return m


def up_matrix(n_in: int, n_out: int) -> np.ndarray:
    """F.interpolate(size=n_out, mode="bilinear", align_corners=False) along one axis (:233, :241): source index
    max(scale (d + 0.5) - 0.5, 0) with scale = n_in / n_out, neighbours clamped at the border."""
    m = np.zeros((n_out, n_in))
    scale = n_in / n_out
    for d in range(n_out):
        src = max(scale * (d + 0.5) - 0.5, 0.0)
        i0 = min(int(np.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        w1 = src - i0
        m[d, i0] += 1.0 - w1
        m[d, i1] += w1
    return m


def torch_quantile(v: np.ndarray, q: float, rank_dtype=np.float32) -> float:
    """torch.quantile(v, q) with the default linear interpolation (:242): ranks = q (n - 1) computed IN THE INPUT'S DTYPE (fp32 for
    the reference's tensors -- at two million pixels that rounds the rank to a multiple of 1/8), lerp between the two order statistics."""
    s = np.sort(np.asarray(v).ravel())
    rank = rank_dtype(rank_dtype(q) * rank_dtype(s.size - 1))
    lo = int(np.floor(rank))
    hi = min(int(np.ceil(rank)), s.size - 1)
    w = float(rank - rank_dtype(lo))
    return float(s[lo] + w * (s[hi] - s[lo]))


def depth_normal_loss(depth: np.ndarray, normal: np.ndarray, tan_fovx: float, tan_fovy: float, scale_factor=None, quantile: float = 0.9,
                      rank_dtype=np.float32, with_grad: bool = False):
    """-> loss, or (loss, dL/ddepth, dL/dnormal) in fp64.  depth (H, W), normal (3, H, W)."""
    d0 = np.asarray(depth, dtype=np.float64)
    nr = np.asarray(normal, dtype=np.float64)
    H0, W0 = d0.shape
    resample = scale_factor is not None and scale_factor != 1
    if resample and scale_factor != 0.5:
        raise ValueError("only scale_factor None, 1 or 0.5 (what the shipped configs use) is restated")
    Dy_, Dx_ = (down_matrix(H0), down_matrix(W0)) if resample else (np.eye(H0), np.eye(W0))
    d = Dy_ @ d0 @ Dx_.T                                                     # :217
    H, W = d.shape
    Uy, Ux = (up_matrix(H, H0), up_matrix(W, W0)) if resample else (np.eye(H0), np.eye(W0))
    gx, gy = correlate3(d, SCHARR_X), correlate3(d, SCHARR_Y)                # :219
    Dx, Dy = gx / d, gy / d                                                  # :220
    x, y = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    cx, cy = x - W / 2 + 0.5, y - H / 2 + 0.5
    n_h = np.stack([W * Dx / (2 * tan_fovx), H * Dy / (2 * tan_fovy), -(1 + cx * Dx + cy * Dy)])   # :226-229
    n_f = np.stack([Uy @ p @ Ux.T for p in n_h])                             # :232-233
    len_f = np.sqrt((n_f * n_f).sum(0))
    dn = n_f / len_f                                                         # :234
    gn_f = Uy @ np.sqrt(gx * gx + gy * gy) @ Ux.T                            # :238-241
    thr = torch_quantile(gn_f, quantile, rank_dtype)                         # :242
    mask = (gn_f < thr).astype(np.float64)                                   # :243
    len_n = np.maximum(np.sqrt((nr * nr).sum(0)), 1e-8)
    nrm = nr / len_n                                                         # :254 (F.normalize, eps = 1e-8)
    N = H0 * W0
    loss = float(((1.0 - (nrm * dn).sum(0)) * mask).sum() / N)               # :255
    if not with_grad:
        return loss
    g_dn = -mask * nrm / N
    g_nrm = -mask * dn / N
    g_normal = (g_nrm - nrm * (nrm * g_nrm).sum(0)) / len_n                  # through x / max(|x|, eps) (|x| > eps)
    g_nf = (g_dn - dn * (dn * g_dn).sum(0)) / len_f
    g_nh = np.stack([Uy.T @ p @ Ux for p in g_nf])
    g_Dx = g_nh[0] * W / (2 * tan_fovx) - g_nh[2] * cx
    g_Dy = g_nh[1] * H / (2 * tan_fovy) - g_nh[2] * cy
    g_gx, g_gy = g_Dx / d, g_Dy / d
    g_dd = -(g_Dx * gx + g_Dy * gy) / (d * d)
    g_dh = correlate3(g_gx, SCHARR_X, transpose=True) + correlate3(g_gy, SCHARR_Y, transpose=True) + g_dd
    g_depth = Dy_.T @ g_dh @ Dx_
    return loss, g_depth, g_normal
