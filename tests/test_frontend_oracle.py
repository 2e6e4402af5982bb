"""CPU: pins oracle/frontend.py (the restatement of the model-side preamble / epilogue) against
torch running the reference's own expressions, copied line by line from src/diff_recon/models/VanillaTS_model.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import frontend as fo  # noqa: E402


def _ulp_diff(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()


@pytest.fixture(scope="module")
def params():
    g = torch.Generator().manual_seed(7)
    P = 5000
    return dict(vertex=torch.randn(P, 3, 3, generator=g) * 3, f_dc=torch.rand(P, 1, 3, generator=g), f_rest=torch.randn(P, 15, 3, generator=g),
                opacity=torch.randn(P, 1, generator=g) * 2, campos=torch.tensor([0.3, -0.2, -4.0]))


def test_gamma_rescale_ratio():
    import scipy.special

    for gamma in (1.0, 2.5, 7.0, 50.0):
        beta = 1 / gamma  # VanillaTS_model.py:616-617
        ref = 1 / np.sqrt(2 ** beta * beta * scipy.special.gamma(beta))
        assert abs(fo.gamma_rescale_ratio(gamma) - ref) <= 1e-15 * ref
    assert abs(fo.gamma_rescale_ratio(1.0) - 1 / np.sqrt(2.0)) < 1e-15


def test_rescale_triangles_matches_torch(params):
    v = params["vertex"]
    for ratio in (fo.gamma_rescale_ratio(7.0), 0.5, 1.7):
        t_center = v.mean(dim=1, keepdim=True)  # :444
        ref = (v - t_center) * ratio + t_center  # :445
        assert np.array_equal(ref.numpy(), fo.rescale_triangles(v.numpy(), ratio, torch_device="cpu"))
        # the CUDA flavour differs only in how the mean is normalised (* fp32(1/3) instead of / 3): at most 1 ulp of the centre
        assert np.abs(ref.numpy() - fo.rescale_triangles(v.numpy(), ratio, torch_device="cuda")).max() <= 2e-6


def test_rescale_backward_is_autograd(params):
    v = params["vertex"].double().requires_grad_(True)
    ratio = fo.gamma_rescale_ratio(3.0)
    g = torch.randn(v.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    t_center = v.mean(dim=1, keepdim=True)
    ((v - t_center) * ratio + t_center).backward(g)
    assert np.allclose(v.grad.numpy(), fo.rescale_triangles_backward(g.numpy(), ratio), rtol=1e-12, atol=1e-14)


def test_opacity_matches_torch(params):
    x = params["opacity"]
    op = torch.sigmoid(x)  # :84
    mine = fo.get_opacity(x.numpy())
    assert _ulp_diff(op.numpy(), mine) <= 2  # torch's CPU sigmoid is a vectorised (Sleef) exp; the CUDA one is held bit-exact on the GPU
    for thr in (0.3, 0.5):
        ste = ((op > thr).float() - op).detach() + op  # :621
        assert np.array_equal(ste.numpy(), fo.opacity_ste(op.numpy(), thr))
        assert set(np.unique(np.round(ste.numpy(), 5))) <= {0.0, 1.0}
    xg = x.double().requires_grad_(True)
    torch.sigmoid(xg).backward(torch.ones_like(xg) * 0.37)
    assert np.allclose(xg.grad.numpy(), fo.sigmoid_backward(np.full(x.shape, 0.37), torch.sigmoid(x.double()).numpy()), rtol=1e-12)


def test_features_and_bg_depth(params):
    ref = torch.cat((params["f_dc"], params["f_rest"]), dim=1)  # :80
    assert np.array_equal(ref.numpy(), fo.get_features(params["f_dc"].numpy(), params["f_rest"].numpy()))
    assert np.array_equal(params["f_dc"].numpy(), fo.get_features(params["f_dc"].numpy(), None))
    bg = (params["campos"].view(1, 1, 3) - params["vertex"]).norm(dim=-1).max()  # :623
    assert abs(float(bg) - fo.bg_depth(params["vertex"].numpy(), params["campos"].numpy())) <= 4e-6 * float(bg)


@pytest.mark.parametrize("s", [1, 2, 3, 4])
def test_bilinear_downsample_matches_interpolate(s):
    g = torch.Generator().manual_seed(s)
    h, w = 9, 14
    x = torch.rand(4, h * s, w * s, generator=g, requires_grad=True)
    y = F.interpolate(x.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0)  # :648
    mine = fo.bilinear_downsample(x.detach().numpy(), s)
    assert mine.shape == (4, h, w)
    assert np.abs(y.detach().numpy() - mine).max() <= 1.2e-7  # CPU torch sums the four taps in a different order for even s
    gg = torch.rand(4, h, w, generator=g)
    y.backward(gg)
    assert np.abs(x.grad.numpy() - fo.bilinear_downsample_backward(gg.numpy(), s)).max() <= 1e-7


def test_training_statistic_matches_reference_lines():
    g = torch.Generator().manual_seed(3)
    P = 4000
    radii = torch.randint(-1, 30, (P,), generator=g).clamp(min=0).int()
    c2d_grad = torch.randn(P, 2, generator=g)
    csum, cmax = torch.rand(P, generator=g), torch.rand(P, generator=g)
    st = {k: torch.rand(P, generator=g) for k in ("gradient_accum", "gradient_denom", "contrib_sum", "contrib_max", "contrib_denom", "max_radii2D")}
    mine = fo.training_statistic({k: v.numpy() for k, v in st.items()}, radii.numpy(), c2d_grad.numpy(), csum.numpy(), cmax.numpy())
    visible_mask = radii > 0  # :674
    st["gradient_accum"][visible_mask] += torch.norm(c2d_grad[visible_mask, :2], dim=-1)  # :358
    st["gradient_denom"][visible_mask] += 1
    st["contrib_sum"][visible_mask] = torch.max(st["contrib_sum"][visible_mask], csum[visible_mask])
    st["contrib_max"][visible_mask] = torch.max(st["contrib_max"][visible_mask], cmax[visible_mask])
    st["contrib_denom"][visible_mask] += 1
    st["max_radii2D"][visible_mask] = torch.max(st["max_radii2D"][visible_mask], radii[visible_mask])  # :363
    for k in st:
        assert np.allclose(st[k].numpy(), mine[k], rtol=1e-6, atol=0), k  # torch.norm vs sqrt(sum sq): a few ulp
