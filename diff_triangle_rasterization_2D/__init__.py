"""Drop-in shim: lets ``from diff_triangle_rasterization_2D import TriangleRasterizationSettings,
TriangleRasterizer`` (src/diff_recon/renderer/triangle_renderer.py:3-6 of the reference) resolve to
the B200-native implementation when this repository root is on ``sys.path``.  Same public names as
the reference package (R2D/diff_triangle_rasterization_2D/__init__.py)."""
from triangle_splatting_b200 import (  # noqa: F401
    TriangleRasterizationSettings,
    TriangleRasterizer,
    _RasterizeTriangles,
    _C,
    debug_run,
)
