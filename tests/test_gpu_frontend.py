"""GPU parity of the parameter-space front-end: TriangleModelRasterizer / model inputs of the C ABI
against the reference's own Python preamble + epilogue (src/diff_recon/models/VanillaTS_model.py:608-656, :347-363) run through
torch on the same device, feeding (a) our reference-shaped op -- itself parity-green against the reference extension -- and
(b) the live reference extension when oracle/_ref is loadable.

Bars: every integer output bit-exact; the activated opacity and the state written from the rescaled vertices bit-equal to what
the torch kernels produce; forward images bit-equal (same kernels, same input bits); gradients within the run-to-run spread of
the fp32 atomics (test_gpu_parity.py explains the bar); the resize epilogue bit-equal to F.interpolate."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import harness
from harness import mismatch_count, rel_err

pytestmark = pytest.mark.gpu

# gradients of two of OUR paths over identical fp32 inputs (same kernels, deterministic sums): equal up to the rounding of the few
# operations that differ between them (fused vs torch sigmoid / rescale backward)
GRAD_MAX, GRAD_FRAC = 1e-3, 1e-3


def _params(sc, dev, seed=11):
    """Raw parameters for a scene: f_dc / f_rest split of its SH tensor, fresh opacity logits."""
    g = torch.Generator().manual_seed(seed)
    s = sc.to(dev)
    logit = (1.5 * torch.randn(sc.P, 1, generator=g)).to(dev)
    return s, dict(vertex=s.vertex.contiguous(), f_dc=s.shs[:, :1, :].contiguous(), f_rest=s.shs[:, 1:, :].contiguous(), opacity=logit)


def _reference_preamble(p, campos, ratio, ste):
    """VanillaTS_model.py:608-623, line by line, on the device the tensors live on."""
    vertex = p["vertex"]
    shs = torch.cat((p["f_dc"], p["f_rest"]), dim=1)  # :80
    opacity = torch.sigmoid(p["opacity"])  # :84
    vertex_in = vertex
    if ratio != 1.0:
        t_center = vertex.mean(dim=1, keepdim=True)  # :444
        vertex_in = (vertex - t_center) * ratio + t_center  # :445
    if ste is not None:
        opacity = ((opacity > ste).float() - opacity).detach() + opacity  # :621
    bg_depth = (campos.view(1, 1, 3) - vertex).norm(dim=-1).max()  # :623
    return vertex_in, shs, opacity, bg_depth


def _settings(s, **over):
    from triangle_splatting_b200 import TriangleRasterizationSettings

    kw = s.settings_kwargs()
    kw.update(over)
    return TriangleRasterizationSettings(**kw)


def _leaf(p):
    return {k: v.clone().requires_grad_(True) for k, v in p.items()}


def _export_model(fwd_state, P, primitive, dev):
    from triangle_splatting_b200 import _lib

    lib = _lib.load()
    op = torch.zeros(P, device=dev)
    bg = torch.zeros(1, device=dev)
    _lib.check(lib.ts2d_export_model(C.c_void_p(fwd_state.data_ptr()), P, _lib.PRIMITIVES[primitive], C.c_void_p(op.data_ptr()),
                                     C.c_void_p(bg.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "export_model")
    torch.cuda.synchronize(dev)
    return op.detach().cpu().numpy(), float(bg.item())


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


@pytest.mark.parametrize("primitive", ["2D", "3D"])
@pytest.mark.parametrize("ratio,ste", [(1.0, None), (0.9839894788945309, 0.3), (0.7071067811865475, None)])
def test_preamble_arithmetic_has_torch_cuda_bits(primitive, ratio, ste, cuda_device):
    """Opacity activation, STE, rescale and background depth done in K1 == the torch CUDA kernels of the preamble."""
    from triangle_splatting_b200 import _C

    dev = cuda_device
    s, p = _params(harness.golden_scene("sh3_rich"), dev)
    vertex_in, shs, opacity, bg_depth = _reference_preamble(p, s.cam["campos"], ratio, ste)
    c = s.cam
    common = (c["image_width"], c["image_height"], c["tanfovx"], c["tanfovy"], c["viewmatrix"], c["projmatrix"], c["campos"], s.sh_degree,
              s.gamma, 1.0)
    unf = _C.rasterize_triangles(*common, float(bg_depth), s.background, vertex_in.contiguous(), shs.contiguous(), torch.Tensor([]),
                                 opacity.contiguous(), False, True, False, primitive=primitive)
    model = _C.ModelInputs(p["f_dc"], p["f_rest"], p["opacity"], ste, ratio, True)
    fus = _C.rasterize_triangles(*common, 0.0, s.background, p["vertex"], None, None, None, False, True, False, primitive=primitive, model=model)
    assert fus[0] == unf[0] and fus[0] > 0
    radii = fus[2].detach().cpu().numpy()
    assert mismatch_count(radii, unf[2].detach().cpu().numpy()) == 0
    vis = radii > 0
    op_k, bg_k = _export_model(fus[7], s.P, primitive, dev)
    assert mismatch_count(_bits(op_k[vis]), _bits(opacity.view(-1).detach().cpu().numpy()[vis])) == 0, "activated opacity bits differ from torch"
    assert abs(bg_k - float(bg_depth)) <= 2e-6 * float(bg_depth)
    # state written from the (rescaled) vertices and the split SH rows -- identical inputs must give identical records
    st_f, st_u = harness.decode_state(s, fus, dev, primitive), harness.decode_state(s, unf, dev, primitive)
    for k in harness.INT_KEYS + harness.STATE_FLOAT_KEYS:
        if k in st_f:
            a, b = np.asarray(st_f[k]), np.asarray(st_u[k])
            a, b = (_bits(a), _bits(b)) if a.dtype == np.float32 else (a, b)
            assert mismatch_count(a, b) == 0, f"state {k}: rescale / SH staging not bit-equal to the torch preamble"
    assert torch.equal(fus[1], unf[1]) and torch.equal(fus[4], unf[4]), "image / normal differ"
    assert rel_err(fus[3].detach().cpu().numpy(), unf[3].detach().cpu().numpy()) <= 1e-6, "depth"


@pytest.mark.parametrize("primitive", ["2D", "3D"])
@pytest.mark.parametrize("name,ratio,ste", [("sh3_rich", 0.9839894788945309, None), ("sh1_M16_dense", 1.0, 0.3), ("sh0_plain", 0.8, 0.5)])
def test_fused_module_equals_reference_preamble_plus_op(primitive, name, ratio, ste, cuda_device):
    """forward + backward: TriangleModelRasterizer(raw parameters) vs the reference lines in torch + the reference-shaped op."""
    from triangle_splatting_b200 import TriangleModelRasterizer, TriangleRasterizer, TriangleRasterizer3D

    dev = cuda_device
    sc = harness.golden_scene(name)
    if sc.shs is None:
        pytest.skip("feature-mode scene")
    s, p0 = _params(sc, dev)
    g_img = s.grads["dL_dout_feature"]

    def loss_of(out):
        loss = (out[0] * g_img).sum()
        if s.rich_info:
            loss = loss + (out[2] * s.grads["dL_dout_depth"]).sum() + (out[3] * s.grads["dL_dout_normal"]).sum()
        return loss

    # reference flow
    p = _leaf(p0)
    vertex_in, shs, opacity, bg_depth = _reference_preamble(p, s.cam["campos"], ratio, ste)
    c2d_a = torch.zeros((s.P, 2), device=dev, requires_grad=True)
    rast = (TriangleRasterizer3D if primitive == "3D" else TriangleRasterizer)(raster_settings=_settings(s, background_depth=float(bg_depth.detach())))
    out_a = rast.forward(vertex=vertex_in, center2D=c2d_a, opacity=opacity, shs=shs, feature=None)
    loss_of(out_a).backward()
    # fused flow
    q = _leaf(p0)
    c2d_b = torch.zeros((s.P, 2), device=dev, requires_grad=True)
    fused = TriangleModelRasterizer(_settings(s), primitive=primitive, ste_threshold=ste, rescale_ratio=ratio)
    out_b = fused.forward(q["vertex"], c2d_b, q["opacity"], q["f_dc"], q["f_rest"] if q["f_rest"].shape[1] else None)
    loss_of(out_b).backward()

    assert len(out_a) == len(out_b)
    assert torch.equal(out_a[1], out_b[1]), "radii"
    assert torch.equal(out_a[0], out_b[0]), "image bits"
    if s.rich_info:
        assert torch.equal(out_a[3], out_b[3]), "normal bits"
        assert rel_err(out_b[2].detach().cpu().numpy(), out_a[2].detach().cpu().numpy()) <= 1e-6, "depth"
        for i in (4, 5):
            assert rel_err(out_b[i].detach().cpu().numpy(), out_a[i].detach().cpu().numpy()) <= 1e-5, "contrib"
    pairs = [("vertex", p["vertex"].grad, q["vertex"].grad), ("f_dc", p["f_dc"].grad, q["f_dc"].grad), ("opacity", p["opacity"].grad, q["opacity"].grad),
             ("center2D", c2d_a.grad, c2d_b.grad)]
    if p0["f_rest"].shape[1]:
        pairs.append(("f_rest", p["f_rest"].grad, q["f_rest"].grad))
    for k, a, b in pairs:
        assert b is not None and a.shape == b.shape, k
        a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
        assert rel_err(b, a) <= GRAD_MAX, f"{k}: rel err {rel_err(b, a):.3e}"
        assert harness.frac_above(b, a, 1e-4, 1e-3) <= GRAD_FRAC, k


@pytest.mark.parametrize("primitive", ["2D", "3D"])
def test_model_inputs_vs_live_reference(primitive, cuda_device):
    """Reference preamble (torch) + the UNMODIFIED reference extension vs the fused entry point, exact arithmetic mode."""
    from triangle_splatting_b200 import _C

    try:
        ref = harness.load_reference(primitive)
    except Exception as ex:  # noqa: BLE001
        pytest.skip(f"reference extension not loadable: {ex}")
    dev = cuda_device
    s, p = _params(harness.golden_scene("sh3_rich"), dev)
    ratio, ste = 0.9839894788945309, 0.3
    vertex_in, shs, opacity, bg_depth = _reference_preamble(p, s.cam["campos"], ratio, ste)
    c = s.cam
    common = (c["image_width"], c["image_height"], c["tanfovx"], c["tanfovy"], c["viewmatrix"], c["projmatrix"], c["campos"], s.sh_degree,
              s.gamma, 1.0)
    r = ref.rasterize_triangles(*common, float(bg_depth), s.background, vertex_in.contiguous(), shs.contiguous(), torch.empty(0, device=dev),
                                opacity.contiguous(), False, True, False)
    old = _C.set_exact(True)
    try:
        model = _C.ModelInputs(p["f_dc"], p["f_rest"], p["opacity"], ste, ratio, True)
        o = _C.rasterize_triangles(*common, 0.0, s.background, p["vertex"], None, None, None, False, True, False, primitive=primitive, model=model)
    finally:
        _C.set_exact(old)
    assert int(r[0]) == int(o[0])
    assert torch.equal(r[2], o[2]), "radii"
    assert torch.equal(r[1], o[1]), "image bits vs the reference"
    assert torch.equal(r[4], o[4]), "normal bits vs the reference"
    assert rel_err(o[3].detach().cpu().numpy(), r[3].detach().cpu().numpy()) <= 1e-6


@pytest.mark.parametrize("s", [1, 2, 3, 4])
def test_downsample_is_interpolate_bilinear(s, cuda_device):
    from triangle_splatting_b200 import bilinear_downsample

    g = torch.Generator().manual_seed(s)
    h, w = 37, 50
    x = torch.rand(5, h * s, w * s, generator=g).to(cuda_device).requires_grad_(True)
    y_ref = F.interpolate(x.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0)  # VanillaTS_model.py:648
    gg = torch.rand(5, h, w, generator=g).to(cuda_device)
    y_ref.backward(gg)
    g_ref = x.grad.clone()
    x.grad = None
    y = bilinear_downsample(x, s)
    if s == 1:
        assert y is x
        return
    assert torch.equal(y, y_ref), f"scale {s}: not bit-equal to F.interpolate"
    y.backward(gg)
    assert torch.allclose(x.grad, g_ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("k,rho_px", [(2, 6.0), (2, 0.5), (3, 0.8)])
def test_render_up_scale_ste_rescale_and_statistics(k, rho_px, cuda_device):
    """The NerfSynthetic *_mesh recipe (config/NerfSynthetic_VanillaTS_mesh.yaml:26-28: ste 0.3, gamma_rescale, up-scale 2, 3D
    rasterizer): fused module vs VanillaTS_model.py:615-656 + :347-363 in torch around the reference-shaped op."""
    from triangle_splatting_b200 import TrainingStatistics, TriangleModelRasterizer, TriangleRasterizer3D, gamma_rescale_ratio
    from triangle_splatting_b200.scenes import make_scene

    dev = cuda_device
    gamma, ste = 7.0, 0.3
    ratio = gamma_rescale_ratio(gamma)
    # rho_px < 1: many triangles whose full-resolution radius is below the up-scale factor -- rendered, but hidden from the
    # statistics by `radii // render_up_scale` (the 3D primitive's radius has no half-pixel floor)
    sc = make_scene("mesh", 30000, 200, 152, sh_degree=0, rich_info=True, gamma=gamma, seed=5, geometry_grads=True, rho_px=rho_px)
    s, p0 = _params(sc, dev)
    w, h = s.cam["image_width"], s.cam["image_height"]
    g = torch.Generator().manual_seed(9)
    g_img, g_dep, g_nrm = [(torch.rand(*shape, generator=g) / (h * w)).to(dev) for shape in ((3, h, w), (h, w), (3, h, w))]
    prev = {f: torch.rand(s.P, generator=g).to(dev) * (30.0 if f == "max_radii2D" else 1.0) for f in TrainingStatistics.FIELDS}

    # ---- reference flow (torch)
    p = _leaf(p0)
    vertex_in, shs, opacity, bg_depth = _reference_preamble(p, s.cam["campos"], ratio, ste)
    c2d_a = torch.zeros((s.P, 2), device=dev, requires_grad=True)
    rast = TriangleRasterizer3D(raster_settings=_settings(s, image_width=w * k, image_height=h * k, background_depth=float(bg_depth.detach())))  # :625-630
    render, radii, depth, normal, csum, cmax = rast.forward(vertex=vertex_in, center2D=c2d_a, opacity=opacity, shs=shs, feature=None)
    render = F.interpolate(render.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0)  # :648
    radii = radii // k  # :651
    depth = F.interpolate(depth.unsqueeze(0).unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0).squeeze(0)  # :653
    normal = F.interpolate(normal.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0)  # :655
    ((render * g_img).sum() + (depth * g_dep).sum() + (normal * g_nrm).sum()).backward()
    st_a = {f: t.clone() for f, t in prev.items()}
    visible_mask = radii > 0  # :674
    st_a["gradient_accum"][visible_mask] += torch.norm(c2d_a.grad[visible_mask, :2], dim=-1)  # :358
    st_a["gradient_denom"][visible_mask] += 1
    st_a["contrib_sum"][visible_mask] = torch.max(st_a["contrib_sum"][visible_mask], csum[visible_mask])
    st_a["contrib_max"][visible_mask] = torch.max(st_a["contrib_max"][visible_mask], cmax[visible_mask])
    st_a["contrib_denom"][visible_mask] += 1
    st_a["max_radii2D"][visible_mask] = torch.max(st_a["max_radii2D"][visible_mask], radii[visible_mask])  # :363

    # ---- fused flow
    q = _leaf(p0)
    c2d_b = torch.zeros((s.P, 2), device=dev, requires_grad=True)
    stats = TrainingStatistics(s.P, dev)
    for f in stats.FIELDS:
        getattr(stats, f).copy_(prev[f])
    fused = TriangleModelRasterizer(_settings(s), primitive="3D", ste_threshold=ste, rescale_ratio=ratio, render_up_scale=k, statistics=stats)
    render_b, radii_b, depth_b, normal_b, csum_b, cmax_b = fused.forward(q["vertex"], c2d_b, q["opacity"], q["f_dc"], None)
    ((render_b * g_img).sum() + (depth_b * g_dep).sum() + (normal_b * g_nrm).sum()).backward()

    assert render_b.shape == (3, h, w) and depth_b.shape == (h, w) and normal_b.shape == (3, h, w)
    assert torch.equal(radii, radii_b)
    if rho_px < 1.0:  # the case this parametrisation exists for must really occur
        full = rast.forward(vertex=vertex_in.detach(), center2D=c2d_a.detach(), opacity=opacity.detach(), shs=shs.detach(), feature=None)[1]
        assert int(((full > 0) & (full // k == 0)).sum()) > 50, "scene has no triangles with 0 < radius < render_up_scale"
    assert torch.equal(render, render_b) and torch.equal(normal, normal_b)
    assert rel_err(depth_b.detach().cpu().numpy(), depth.detach().cpu().numpy()) <= 1e-6
    # radii // k can hide triangles whose radius is below the up-scale factor from `visible_mask`; the statistics follow the
    # reference's mask (radii // k > 0), the gradients do not depend on it
    for f in stats.FIELDS:
        a, b = st_a[f].detach().cpu().numpy(), getattr(stats, f).detach().cpu().numpy()
        if f == "gradient_accum":
            assert rel_err(b, a) <= 2e-2 and harness.frac_above(b, a, 1e-4, 1e-3) <= GRAD_FRAC, f
        else:
            assert np.allclose(b, a, rtol=1e-5, atol=0), f
    for kk, a, b in (("vertex", p["vertex"].grad, q["vertex"].grad), ("f_dc", p["f_dc"].grad, q["f_dc"].grad),
                     ("opacity", p["opacity"].grad, q["opacity"].grad), ("center2D", c2d_a.grad, c2d_b.grad)):
        a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
        assert rel_err(b, a) <= GRAD_MAX, f"{kk}: {rel_err(b, a):.3e}"
        assert harness.frac_above(b, a, 1e-4, 1e-3) <= GRAD_FRAC, kk


def test_model_input_errors_and_empty(cuda_device):
    from triangle_splatting_b200 import TriangleModelRasterizer

    dev = cuda_device
    s, p = _params(harness.golden_scene("sh3_rich"), dev)
    c2d = torch.zeros((s.P, 2), device=dev)
    r = TriangleModelRasterizer(_settings(s))
    with pytest.raises(RuntimeError):
        r.forward(p["vertex"], c2d, p["opacity"], p["f_dc"][:, 0, :], p["f_rest"])  # f_dc must be (P, 1, 3)
    with pytest.raises(RuntimeError):
        r.forward(p["vertex"], c2d, p["opacity"][:-1], p["f_dc"], p["f_rest"])
    with pytest.raises(RuntimeError):
        r.forward(p["vertex"].cpu(), c2d, p["opacity"], p["f_dc"], p["f_rest"])  # no CPU path
    with pytest.raises(RuntimeError):
        TriangleModelRasterizer(_settings(s), rescale_ratio=0.0).forward(p["vertex"], c2d, p["opacity"], p["f_dc"], p["f_rest"])
    with pytest.raises(ValueError):
        TriangleModelRasterizer(_settings(s), render_up_scale=0)
    e = lambda *shape: torch.zeros(shape, device=dev)
    out = r.forward(e(0, 3, 3), e(0, 2), e(0, 1), e(0, 1, 3), e(0, 15, 3))
    assert out[0].shape == (3, s.cam["image_height"], s.cam["image_width"]) and out[1].numel() == 0
    # a warp-ragged triangle count (P % 32 != 0) with f_rest rows that are not 16-byte multiples (M = 4 -> 36 B rows)
    n = 1000 + 13
    out = TriangleModelRasterizer(_settings(s, sh_degree=1)).forward(p["vertex"][:n].contiguous(), c2d[:n], p["opacity"][:n].contiguous(),
                                                                      p["f_dc"][:n].contiguous(), p["f_rest"][:n, :3].contiguous())
    ref_shs = torch.cat((p["f_dc"][:n], p["f_rest"][:n, :3]), dim=1).contiguous()
    from triangle_splatting_b200 import TriangleRasterizer

    bg = (s.cam["campos"].view(1, 1, 3) - p["vertex"][:n]).norm(dim=-1).max()
    ref = TriangleRasterizer(raster_settings=_settings(s, sh_degree=1, background_depth=float(bg.detach()))).forward(
        vertex=p["vertex"][:n].contiguous(), center2D=c2d[:n], opacity=torch.sigmoid(p["opacity"][:n]), shs=ref_shs, feature=None)
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
