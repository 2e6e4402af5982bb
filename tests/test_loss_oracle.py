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


# ---- depth-normal consistency loss (the geometry term) ---------------------------------------------------------------------------
GOLDEN = os.path.join(ROOT, "tests", "golden")
DN_CASES = ["half_38x50", "half_odd_37x53", "full_21x25", "half_96x128"]


def _load_dn(name):
    z = np.load(os.path.join(GOLDEN, f"depth_normal_{name}.npz"))
    sf = float(z["scale_factor"])
    return z, (None if sf < 0 else sf)


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(float(np.abs(b).max()), 1e-30))


@pytest.mark.parametrize("name", DN_CASES)
def test_depth_normal_oracle_vs_reference_fixtures(name):
    """oracle/loss.py: depth_normal_loss against the outputs of the reference's own DepthNormalLoss (tests/golden/make_loss_golden.py,
    fp32 on the CPU): loss to 1e-6, both gradients to 2e-5 of their largest entry (the fixtures are fp32, the oracle fp64)."""
    z, sf = _load_dn(name)
    loss, g_depth, g_normal = lo.depth_normal_loss(z["depth"], z["normal"], float(z["tan_fovx"]), float(z["tan_fovy"]), sf, with_grad=True)
    assert abs(loss - float(z["loss"])) <= 1e-6 * abs(float(z["loss"]))
    assert _rel(g_depth, z["g_depth"]) <= 2e-5
    assert _rel(g_normal, z["g_normal"]) <= 2e-5


def test_depth_normal_oracle_vs_live_reference_class():
    """Same comparison against the class itself when the reference tree is present (it is in the build container, not on the GPU box)."""
    sys.path.insert(0, GOLDEN)
    import make_loss_golden as mk

    if not os.path.exists(mk.REF):
        pytest.skip("/root/reference not present")
    ref = mk.load_reference_module()
    for h, w, sf in ((30, 44, 0.5), (33, 41, 0.5), (26, 26, None), (24, 40, 1)):
        depth, normal = mk.scene(h, w, seed=7 * h + w)
        d, n = depth.clone().requires_grad_(True), normal.clone().requires_grad_(True)
        loss = ref.DepthNormalLoss(scale_factor=sf)(d, n, 0.6, 0.45)
        loss.backward()
        o = lo.depth_normal_loss(depth.numpy(), normal.numpy(), 0.6, 0.45, sf, with_grad=True)
        assert abs(o[0] - float(loss)) <= 1e-6 * abs(float(loss)), (h, w, sf)
        assert _rel(o[1], d.grad.numpy()) <= 2e-5 and _rel(o[2], n.grad.numpy()) <= 2e-5, (h, w, sf)
        # depth_grad / normal_grad = False detach the respective input (:250-253): the other gradient is unchanged
        d2, n2 = depth.clone().requires_grad_(True), normal.clone().requires_grad_(True)
        ref.DepthNormalLoss(scale_factor=sf, depth_grad=False)(d2, n2, 0.6, 0.45).backward()
        assert d2.grad is None and torch.equal(n2.grad, n.grad)


def test_depth_normal_oracle_gradient_is_the_derivative():
    """Central differences in fp64 on the oracle itself (mask held: perturbations far below the distance of any pixel to the threshold)."""
    z, sf = _load_dn("half_odd_37x53")
    depth, normal = z["depth"].astype(np.float64), z["normal"].astype(np.float64)
    args = (float(z["tan_fovx"]), float(z["tan_fovy"]), sf)
    _, g_depth, g_normal = lo.depth_normal_loss(depth, normal, *args, rank_dtype=np.float64, with_grad=True)
    rng = np.random.default_rng(0)
    for _ in range(12):
        y, x = rng.integers(0, depth.shape[0]), rng.integers(0, depth.shape[1])
        e = 1e-6
        dp, dm = depth.copy(), depth.copy()
        dp[y, x] += e
        dm[y, x] -= e
        fd = (lo.depth_normal_loss(dp, normal, *args, rank_dtype=np.float64) - lo.depth_normal_loss(dm, normal, *args, rank_dtype=np.float64)) / (2 * e)
        assert abs(fd - g_depth[y, x]) <= 1e-6 * max(abs(g_depth).max(), 1e-12) + 1e-4 * abs(g_depth[y, x]), (y, x, fd, g_depth[y, x])
        c = rng.integers(0, 3)
        np_, nm = normal.copy(), normal.copy()
        np_[c, y, x] += e
        nm[c, y, x] -= e
        fd = (lo.depth_normal_loss(depth, np_, *args, rank_dtype=np.float64) - lo.depth_normal_loss(depth, nm, *args, rank_dtype=np.float64)) / (2 * e)
        assert abs(fd - g_normal[c, y, x]) <= 1e-6 * max(abs(g_normal).max(), 1e-12) + 1e-4 * abs(g_normal[c, y, x])


def test_torch_quantile_restatement():
    g = torch.Generator().manual_seed(5)
    for n in (7, 100, 4097):
        v = torch.rand(n, generator=g)
        for q in (0.0, 0.3, 0.9, 1.0):
            assert abs(lo.torch_quantile(v.numpy(), q) - float(torch.quantile(v, q))) <= 1e-7
