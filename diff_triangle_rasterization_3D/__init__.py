"""Drop-in shim: lets ``from diff_triangle_rasterization_3D import TriangleRasterizationSettings,
TriangleRasterizer`` (src/diff_recon/renderer/triangle_renderer.py:32-36 of the reference, taken when
``rasterizer_type == "3D"`` -- the shipped ``*_mesh`` configs) resolve to the B200-native implementation
when this repository root is on ``sys.path``.  Same public names as the reference package
(R3D/diff_triangle_rasterization_3D/__init__.py)."""
from triangle_splatting_b200 import (  # noqa: F401
    TriangleRasterizationSettings,
    _C,
    debug_run,
)
from triangle_splatting_b200 import TriangleRasterizer3D as TriangleRasterizer  # noqa: F401
from triangle_splatting_b200 import _RasterizeTriangles3D as _RasterizeTriangles  # noqa: F401
