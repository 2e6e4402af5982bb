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


def test_cpu_tensors_fail_loudly():
    """There is no CPU fallback: CPU inputs must raise, not silently compute."""
    sc = make_scene("x", 4, 32, 32)
    args = harness._fwd_args(sc)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _C.rasterize_triangles(*args)


def test_argument_errors_precede_device_work():
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
