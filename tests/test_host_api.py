"""CPU tests of the Python host mirror of the reference interface (no GPU work)."""
import pytest
import torch

import harness  # noqa: F401
from triangle_splatting_b200 import TriangleRasterizationSettings, TriangleRasterizer, _C
from triangle_splatting_b200.scenes import make_camera, make_scene

REFERENCE_FIELDS = ("image_width", "image_height", "tanfovx", "tanfovy", "viewmatrix", "projmatrix", "campos", "sh_degree", "gamma",
                    "scale_modifier", "background_depth", "background", "back_culling", "rich_info", "debug")


def test_settings_fields_match_reference_order():
    # R2D/diff_triangle_rasterization_2D/__init__.py:28-46
    assert TriangleRasterizationSettings._fields == REFERENCE_FIELDS


def test_shim_exports_reference_names():
    import diff_triangle_rasterization_2D as shim

    for n in ("TriangleRasterizationSettings", "TriangleRasterizer", "_RasterizeTriangles", "_C"):
        assert hasattr(shim, n)
    assert hasattr(shim._C, "rasterize_triangles") and hasattr(shim._C, "rasterize_triangles_backward")


def test_rasterizer_requires_exactly_one_colour_source():
    sc = make_scene("x", 4, 32, 32)
    r = TriangleRasterizer(TriangleRasterizationSettings(**sc.settings_kwargs()))
    assert r.raster_settings.image_width == 32
    with pytest.raises(Exception, match="excatly one"):
        r(vertex=sc.vertex, center2D=None, opacity=sc.opacity)
    with pytest.raises(Exception, match="excatly one"):
        r(vertex=sc.vertex, center2D=None, opacity=sc.opacity, shs=sc.shs, feature=torch.zeros(4, 3))


def test_derive_matches_reference_rules():
    v = torch.zeros(5, 3, 3)
    assert _C._derive(v, torch.zeros(5, 16, 3), torch.Tensor([])) == (5, True, 3, 16)
    assert _C._derive(v, torch.Tensor([]), torch.zeros(5, 2)) == (5, False, 2, 0)
    assert _C._derive(v, torch.zeros(5, 4, 3), torch.zeros(0, 3)) == (5, True, 3, 4)


@pytest.fixture(params=["native", "ctypes"])
def host_layer(request, monkeypatch):
    """Both host layers over the C ABI: the compiled pybind11 module (csrc/ts2d_pybind.cpp) and the ctypes code of _C.py."""
    if request.param == "native":
        assert _C.native() is not None, "the compiled host layer must be built (python -m triangle_splatting_b200.build)"
    else:
        monkeypatch.setattr(_C, "_NATIVE", False)
    return request.param


def test_compiled_host_layer_mirrors_the_reference_pybind_module():
    """R2D/ext.cpp:4-9: a pybind11 module with rasterize_triangles / rasterize_triangles_backward, here on top of the C ABI."""
    from triangle_splatting_b200 import _lib

    nat = _C.native()
    assert nat is not None and nat.__name__.endswith("_C_native")
    assert nat.abi_version() == _lib.ABI_VERSION
    for n in ("rasterize_triangles", "rasterize_triangles_backward", "FrameCounters", "configure", "forget_shapes"):
        assert hasattr(nat, n)
    doc = nat.rasterize_triangles.__doc__
    # the reference's positional order (extension_interface.cu:19-40)
    order = ["image_width", "image_height", "tan_fovx", "tan_fovy", "viewmatrix", "projmatrix", "campos", "sh_degree", "gamma", "scale_modifier",
             "background_depth", "background", "vertex", "shs", "feature", "opacity", "back_culling", "rich_info", "debug"]
    pos = [doc.index(n + ":") for n in order]
    assert pos == sorted(pos)
    old = _C.configure()
    assert _C.configure(sync_forward=True)[0] == old[0] and nat.configure()[0] is True
    _C.configure(*old)
    assert tuple(nat.configure()) == tuple(old)


def test_cpu_tensors_fail_loudly(host_layer):
    """There is no CPU fallback: CPU inputs must raise, not silently compute."""
    sc = make_scene("x", 4, 32, 32)
    args = harness._fwd_args(sc)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.rasterize_triangles(*args)


def test_argument_errors_precede_device_work(host_layer):
    sc = make_scene("x", 4, 32, 32)
    args = list(harness._fwd_args(sc))
    args[12] = sc.vertex.reshape(4, 9)
    with pytest.raises(RuntimeError, match="vertex must have dimensions"):
        _C.rasterize_triangles(*args)
    args = list(harness._fwd_args(sc))
    args[8] = -0.5
    with pytest.raises(RuntimeError, match="gamma must be larger than 0"):
        _C.rasterize_triangles(*args)
    args = list(harness._fwd_args(sc))
    args[14] = torch.zeros(4, 5)
    args[13] = torch.Tensor([])
    with pytest.raises(RuntimeError, match="MAX_CHANNELS"):
        _C.rasterize_triangles(*args)


def test_camera_follows_reference_conventions():
    cam = make_camera(640, 480)
    view, full = cam["viewmatrix"], cam["projmatrix"]
    # row-vector convention: p_view = [p,1] @ view ; camera at (0,0,-4) looking down +z
    p = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ view
    assert torch.allclose(p[:3], torch.tensor([0.0, 0.0, 4.0]))
    assert torch.allclose(cam["campos"], torch.tensor([0.0, 0.0, -4.0]))
    h = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ full
    assert h[3] == pytest.approx(4.0) and 0 < float(h[2] / h[3]) < 1
    assert cam["tanfovx"] == pytest.approx(1 / 2.4)


def test_front_end_and_loss_host_checks_without_a_gpu():
    """Parameter-space front-end and fused image loss: constructor / argument checks and the no-CPU-path rule (nothing reaches the device)."""
    import math

    from triangle_splatting_b200 import ImageLoss, TrainingStatistics, TriangleModelRasterizer, bilinear_downsample, gamma_rescale_ratio, image_loss

    sc = make_scene("x", 8, 32, 32, sh_degree=1)
    s = TriangleRasterizationSettings(**sc.settings_kwargs())
    assert abs(gamma_rescale_ratio(1.0) - 1 / math.sqrt(2.0)) < 1e-15  # VanillaTS_model.py:616-617 at gamma = 1: 1 / sqrt(2 * 1 * Gamma(1))
    with pytest.raises(ValueError, match="rasterizer type"):
        TriangleModelRasterizer(s, primitive="4D")  # triangle_renderer.py:35-36
    for bad in (0, 1.5):
        with pytest.raises(ValueError):
            TriangleModelRasterizer(s, render_up_scale=bad)
    r = TriangleModelRasterizer(s, ste_threshold=0.3, rescale_ratio=0.9, render_up_scale=2)
    assert r.raster_settings is s and r.render_up_scale == 2
    f_dc, f_rest = sc.shs[:, :1].contiguous(), sc.shs[:, 1:].contiguous()
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        r.forward(sc.vertex, torch.zeros(sc.P, 2), torch.zeros(sc.P, 1), f_dc, f_rest)
    st = TrainingStatistics(5, "cpu")
    assert tuple(st.as_dict()) == _C.STAT_FIELDS and all(t.shape == (5,) for t in st.as_dict().values())
    x = torch.rand(3, 8, 8)
    assert bilinear_downsample(x, 1) is x
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        bilinear_downsample(x, 2)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        image_loss(x, x.clone(), 0.2)
    assert ImageLoss(0.2).w_ssim == 0.2
    # model inputs are validated before any device work
    mi = _C.ModelInputs(f_dc[:, 0], f_rest, torch.zeros(sc.P, 1))
    with pytest.raises(RuntimeError, match="f_dc must have dimensions"):
        mi.check(sc.P)
    with pytest.raises(RuntimeError, match="rescale_ratio"):
        _C.ModelInputs(f_dc, f_rest, torch.zeros(sc.P, 1), rescale_ratio=0.0).check(sc.P)
    assert _C.ModelInputs(f_dc, f_rest, torch.zeros(sc.P, 1)).M == 4 and _C.ModelInputs(f_dc, None, torch.zeros(sc.P, 1)).M == 1
