"""CPU checks of the drop-in boundary: libts2d.so loads without a GPU and exports every symbol
include/ts2d.h declares; the ctypes binding table covers exactly that set.  No compute calls."""
import os
import re

import harness  # noqa: F401  (sys.path)
from triangle_splatting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ts2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(ts2d_[a-z0-9_]+)\s*\(", src))


def test_header_and_binding_agree():
    decl = _declared()
    assert decl, "no declarations parsed"
    assert decl == set(_lib.SYMBOLS), (decl ^ set(_lib.SYMBOLS))


def test_library_loads_and_exports_everything():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.ts2d_abi_version() == _lib.ABI_VERSION == 4
    assert b"vertex must have dimensions" in lib.ts2d_error_string(-1)
    assert lib.ts2d_error_string(0) == b"ok"


def test_ctypes_structs_match_the_header_layout():
    """The ctypes mirrors must have the C layout of include/ts2d.h (checked by compiling the header with gcc)."""
    import ctypes
    import subprocess
    import tempfile

    names = {"ts2d_camera": _lib.Camera, "ts2d_geometry": _lib.Geometry, "ts2d_flags": _lib.Flags, "ts2d_forward_out": _lib.ForwardOut,
             "ts2d_loss_in": _lib.LossIn, "ts2d_backward_out": _lib.BackwardOut, "ts2d_model_inputs": _lib.ModelInputs,
             "ts2d_model_grads": _lib.ModelGrads, "ts2d_fabric": _lib.FabricC}
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(f'#include <stdio.h>\n#include "ts2d.h"\nint main(void){{{body} return 0;}}')
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout
    sizes = dict(ln.split() for ln in out.strip().splitlines())
    for n, cls in names.items():
        assert int(sizes[n]) == ctypes.sizeof(cls), n


def test_state_sizes_are_monotone_and_aligned():
    lib = _lib.load()
    a, b = lib.ts2d_geometry_state_bytes(1000), lib.ts2d_geometry_state_bytes(2000)
    assert 0 < a < b and a % 256 == 0
    assert lib.ts2d_image_state_bytes(1920, 1080) >= 1920 * 1080 * 8
    assert lib.ts2d_binning_state_bytes(10**6, 1920, 1080) >= 16 * 10**6
    assert lib.ts2d_backward_scratch_bytes(1000) >= 64 * 1000


def test_no_oracle_in_product_path():
    """The product package must never import or link the oracle."""
    pkg = os.path.join(ROOT, "triangle_splatting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU or PyTorch fallback", ""), f"{f} mentions the oracle"
