"""CPU restatement of the model-side preamble / epilogue around the rasterizer call (numpy).

TEST INFRASTRUCTURE ONLY -- like everything under oracle/: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product path never does.

Each function follows the reference's Python (src/diff_recon/models/VanillaTS_model.py, cited per function) in fp32 with the
operation order of the torch kernels it launches.  Pinning: tests/test_frontend_oracle.py runs the same expressions,
copied line by line from the reference, through torch on the CPU and requires these restatements to reproduce them
(bit-exactly for the elementwise ones).  The reference ships no tests or golden vectors for this code either.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def gamma_rescale_ratio(gamma: float) -> float:
    """VanillaTS_model.py:616-617 -- beta = 1 / gamma; rescale_ratio = 1 / np.sqrt(2**beta * beta * scipy.special.gamma(beta))."""
    beta = 1.0 / float(gamma)
    return 1.0 / math.sqrt(2.0 ** beta * beta * math.gamma(beta))


def get_opacity(opacity_logit: np.ndarray) -> np.ndarray:
    """:84 -- torch.sigmoid(self._opacity): 1 / (1 + exp(-x)) evaluated in fp32."""
    x = np.asarray(opacity_logit, dtype=F)
    return (F(1) / (F(1) + np.exp(-x, dtype=F))).astype(F)


def opacity_ste(opacity: np.ndarray, threshold: float) -> np.ndarray:
    """:620-621 -- ((opacity > ste_threshold).float() - opacity).detach() + opacity (forward value)."""
    o = np.asarray(opacity, dtype=F)
    return (((o > F(threshold)).astype(F) - o).astype(F) + o).astype(F)


def get_features(f_dc: np.ndarray, f_rest: np.ndarray | None) -> np.ndarray:
    """:80 -- torch.cat((self._f_dc, self._f_rest), dim=1)."""
    if f_rest is None or f_rest.size == 0:
        return np.ascontiguousarray(f_dc, dtype=F)
    return np.ascontiguousarray(np.concatenate([f_dc, f_rest], axis=1), dtype=F)


def rescale_triangles(vertex: np.ndarray, ratio: float, torch_device: str = "cuda") -> np.ndarray:
    """:431-447 -- t_center = vertex.mean(dim=1, keepdim=True); (vertex - t_center) * rescale_ratio + t_center.
    torch's mean sums the three vertices in order; its CUDA reduce kernel (MeanOps) then multiplies by fp32(1/3) while the CPU
    path is sum().div_(3) -- the reference trains on CUDA, so "cuda" is the flavour the kernels are held to and "cpu" exists
    to pin this restatement against torch in a GPU-less container.  The Python float ratio is cast to fp32 by the mul kernel."""
    v = np.asarray(vertex, dtype=F)
    ssum = ((v[:, 0] + v[:, 1]).astype(F) + v[:, 2]).astype(F)
    c = ((ssum * (F(1) / F(3))) if torch_device == "cuda" else (ssum / F(3))).astype(F)[:, None, :]
    return (((v - c).astype(F) * F(ratio)).astype(F) + c).astype(F)


def rescale_triangles_backward(grad_rescaled: np.ndarray, ratio: float) -> np.ndarray:
    """adjoint of the map above (what autograd produces through :444-446): r * g_i + (1 - r) / 3 * sum_j g_j (fp64 truth)."""
    g = np.asarray(grad_rescaled, dtype=np.float64)
    return ratio * g + (1.0 - ratio) / 3.0 * g.sum(axis=1, keepdims=True)


def sigmoid_backward(grad_opacity: np.ndarray, opacity: np.ndarray) -> np.ndarray:
    """SigmoidBackward: grad * (1 - y) * y."""
    y = np.asarray(opacity, dtype=np.float64).reshape(grad_opacity.shape)
    return np.asarray(grad_opacity, dtype=np.float64) * (1.0 - y) * y


def bg_depth(vertex: np.ndarray, campos: np.ndarray) -> float:
    """:623 -- (camera.camera_center.view(1, 1, 3) - vertex).norm(dim=-1).max(), over the un-rescaled vertices."""
    d = (np.asarray(campos, dtype=F).reshape(1, 1, 3) - np.asarray(vertex, dtype=F)).astype(F)
    return float(np.sqrt((d * d).sum(axis=-1, dtype=F), dtype=F).max())


def _taps(n_out: int, s: int):
    """area_pixel_compute_source_index(scale = s, dst, align_corners=False): src = s * (d + 0.5) - 0.5 (>= 0 for s >= 1)."""
    d = np.arange(n_out, dtype=np.float64)
    src = s * (d + 0.5) - 0.5
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, n_out * s - 1)
    l1 = (src - i0).astype(F)
    return i0, i1, (F(1) - l1).astype(F), l1


def bilinear_downsample(x: np.ndarray, s: int) -> np.ndarray:
    """:647-655 -- F.interpolate(x.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0) for x of shape (planes, h*s, w*s):
    w_y0 * (w_x0 * a + w_x1 * b) + w_y1 * (w_x0 * c + w_x1 * d) in fp32 (upsample_bilinear2d, align_corners=False)."""
    x = np.asarray(x, dtype=F)
    _, hi, wi = x.shape
    h, w = hi // s, wi // s
    y0, y1, wy0, wy1 = _taps(h, s)
    x0, x1, wx0, wx1 = _taps(w, s)
    a, b = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    c, d = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    top = ((wx0 * a).astype(F) + (wx1 * b).astype(F)).astype(F)
    bot = ((wx0 * c).astype(F) + (wx1 * d).astype(F)).astype(F)
    return ((wy0[None, :, None] * top).astype(F) + (wy1[None, :, None] * bot).astype(F)).astype(F)


def bilinear_downsample_backward(g: np.ndarray, s: int) -> np.ndarray:
    """adjoint of bilinear_downsample (fp64 accumulate)."""
    g = np.asarray(g, dtype=np.float64)
    p, h, w = g.shape
    out = np.zeros((p, h * s, w * s), dtype=np.float64)
    y0, y1, wy0, wy1 = _taps(h, s)
    x0, x1, wx0, wx1 = _taps(w, s)
    for (yy, wy) in ((y0, wy0), (y1, wy1)):
        for (xx, wx) in ((x0, wx0), (x1, wx1)):
            np.add.at(out, (slice(None), yy[:, None], xx[None, :]), g * (wy.astype(np.float64)[:, None] * wx.astype(np.float64)[None, :]))
    return out


def training_statistic(stats: dict, radii: np.ndarray, center2D_grad: np.ndarray, contrib_sum: np.ndarray, contrib_max: np.ndarray) -> dict:
    """:347-363 -- _training_statistic on numpy copies of the six accumulators (returns the updated dict)."""
    out = {k: np.array(v, dtype=F, copy=True) for k, v in stats.items()}
    vis = np.asarray(radii) > 0  # :674 visible_mask
    out["gradient_accum"][vis] += np.sqrt((np.asarray(center2D_grad, dtype=F)[vis, :2] ** 2).sum(axis=-1, dtype=F), dtype=F)
    out["gradient_denom"][vis] += 1
    out["contrib_sum"][vis] = np.maximum(out["contrib_sum"][vis], np.asarray(contrib_sum, dtype=F)[vis])
    out["contrib_max"][vis] = np.maximum(out["contrib_max"][vis], np.asarray(contrib_max, dtype=F)[vis])
    out["contrib_denom"][vis] += 1
    out["max_radii2D"][vis] = np.maximum(out["max_radii2D"][vis], np.asarray(radii)[vis].astype(F))
    return out
