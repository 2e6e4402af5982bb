"""The reference's OWN caller, unchanged, on top of the import-name shims (CPU; needs the reference tree, which exists in the build
container only -- skipped elsewhere).

`src/diff_recon/renderer/triangle_renderer.py` is imported from /root/reference as it lies (never copied): its
`from diff_triangle_rasterization_2D import ...` / `..._3D import ...` lines (:3-10) resolve to this repository's shim packages, its
`TriangleRenderer.__init__` builds the settings tuple with the reference's keywords (:38-54) and `render()` calls
`rasterizer.forward(vertex=, center2D=, opacity=, shs=, feature=)` (:69-75).  There is no GPU here, so the call must travel through
the autograd.Function and the host layer, pass every shape / argument check, and stop at the one thing a CPU box cannot satisfy:
"must be a CUDA tensor" -- raised by OUR extension boundary (there is no CPU path to fall back to).  The camera comes from the
reference's own `Camera` class (utils/camera.py:72-116).

The parent packages `diff_recon`, `diff_recon.renderer`, `diff_recon.utils` are entered into sys.modules as bare namespaces pointing
at the reference directories: the real `diff_recon/__init__.py` imports every trainer and dataset (torchmetrics, open3d, ... -- not in
this image and out of scope), the two modules under test import none of that.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

import harness  # noqa: F401  (puts the repository root on sys.path)

REF_SRC = os.path.join(os.environ.get("TS2D_REFERENCE_ROOT", "/root/reference"), "src", "diff_recon")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree is not present on this box")


@pytest.fixture()
def reference_modules():
    names = {"diff_recon": REF_SRC, "diff_recon.renderer": os.path.join(REF_SRC, "renderer"), "diff_recon.utils": os.path.join(REF_SRC, "utils")}
    saved = {n: sys.modules.get(n) for n in list(names) + ["diff_recon.renderer.triangle_renderer", "diff_recon.utils.camera"]}
    for n, path in names.items():
        m = types.ModuleType(n)
        m.__path__ = [path]
        sys.modules[n] = m
    try:
        tr = importlib.import_module("diff_recon.renderer.triangle_renderer")
        cam = importlib.import_module("diff_recon.utils.camera")
        yield tr, cam
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_reference_renderer_binds_to_the_shims(reference_modules):
    tr, _ = reference_modules
    import triangle_splatting_b200 as ours

    assert os.path.realpath(tr.__file__).startswith(os.path.realpath(REF_SRC))
    assert tr.TriangleRasterizer_2D is ours.TriangleRasterizer and tr.TriangleRasterizationSettings_2D is ours.TriangleRasterizationSettings
    assert tr.TriangleRasterizer_3D is ours.TriangleRasterizer3D and tr.TriangleRasterizationSettings_3D is ours.TriangleRasterizationSettings


@pytest.mark.parametrize("rasterizer_type", ["2D", "3D"])
@pytest.mark.parametrize("rich_info", [False, True])
def test_unchanged_caller_reaches_our_extension(reference_modules, rasterizer_type, rich_info):
    tr, camera = reference_modules
    from triangle_splatting_b200 import _C

    cam = camera.Camera(R=np.eye(3), T=np.array([0.0, 0.0, 4.0]), FoVx=1.0, FoVy=0.8, image_width=96, image_height=64)
    renderer = tr.TriangleRenderer(cam, bg_depth=50.0, bg_color=torch.Tensor([0, 0, 0]), sh_degree=3, gamma=1.0, rich_info=rich_info,
                                   rasterizer_type=rasterizer_type)
    s = renderer.rasterizer.raster_settings
    assert (s.image_width, s.image_height, s.sh_degree, s.rich_info) == (96, 64, 3, rich_info)
    assert s.viewmatrix is cam.world_view_transform and s.projmatrix is cam.full_proj_transform and s.campos is cam.camera_center
    g = torch.Generator().manual_seed(0)
    P = 7
    vertex = torch.randn(P, 3, 3, generator=g, requires_grad=True)
    shs = torch.randn(P, 16, 3, generator=g, requires_grad=True)
    opacity = torch.rand(P, 1, generator=g, requires_grad=True)
    seen = {}
    real = _C.rasterize_triangles

    def spy(*args, **kw):  # what the unchanged caller hands to the extension boundary
        seen["args"], seen["kw"] = args, kw
        return real(*args, **kw)

    _C.rasterize_triangles = spy
    try:
        with pytest.raises(RuntimeError, match="must be a CUDA tensor: this rasterizer has no CPU path"):
            renderer.render(vertex=vertex, shs=shs, color=None, opacity=opacity)
    finally:
        _C.rasterize_triangles = real
    a = seen["args"]
    assert len(a) == 19  # the reference's positional signature, extension_interface.cu:19-40
    assert a[0] == 96 and a[1] == 64 and a[7] == 3 and a[10] == 50.0 and a[17] is rich_info
    assert a[12] is vertex and a[13] is shs and a[14].numel() == 0 and a[15] is opacity
    assert seen["kw"]["primitive"] == rasterizer_type
