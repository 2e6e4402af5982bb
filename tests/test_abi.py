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
    assert lib.ts2d_abi_version() == _lib.ABI_VERSION == 7
    assert b"vertex must have dimensions" in lib.ts2d_error_string(-1)
    assert lib.ts2d_error_string(0) == b"ok"


def test_ctypes_structs_match_the_header_layout():
    """The ctypes mirrors must have the C layout of include/ts2d.h (checked by compiling the header with gcc)."""
    import ctypes
    import subprocess
    import tempfile

    names = {"ts2d_camera": _lib.Camera, "ts2d_geometry": _lib.Geometry, "ts2d_flags": _lib.Flags, "ts2d_forward_out": _lib.ForwardOut,
             "ts2d_loss_in": _lib.LossIn, "ts2d_backward_out": _lib.BackwardOut, "ts2d_model_inputs": _lib.ModelInputs,
             "ts2d_model_grads": _lib.ModelGrads, "ts2d_frame_counters": _lib.FrameCounters}
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
    # capacity is the inverse of the size function (forward and backward derive the array layout from the blob size alone)
    for cap in (1, 2, 3071, 3072, 3073, 10**6, 6946149, 23 * 10**6):
        b = lib.ts2d_binning_state_bytes(cap, 1920, 1080)
        got = lib.ts2d_binning_capacity(b)
        assert got >= cap and lib.ts2d_binning_state_bytes(got, 1920, 1080) <= b, (cap, got)
        assert lib.ts2d_binning_capacity(b - 256) < got
    assert lib.ts2d_binning_capacity(0) == 0
    bb = lib.ts2d_binning_state_bytes(5000, 64, 64)
    assert lib.ts2d_backward_scratch_bytes(1000, bb, 0) >= 64 * 1000
    assert lib.ts2d_backward_scratch_bytes(1000, bb, 300) >= lib.ts2d_backward_scratch_bytes(1000, bb, 0) + 64 * 290


def test_no_oracle_in_product_path():
    """The product package must never import or link the oracle."""
    pkg = os.path.join(ROOT, "triangle_splatting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU or PyTorch fallback", ""), f"{f} mentions the oracle"


def _dummy_call_structs(P=4, **geom_over):
    """Camera / geometry / flags with non-NULL (never dereferenced) pointers: argument validation happens before any device work."""
    import ctypes as C

    fake = C.c_void_p(0x1000)
    cam = _lib.Camera(64, 64, 0.5, 0.5, fake, fake, fake)
    g = dict(P=P, sh_degree=0, M=1, C=3, use_shs=1, gamma=1.0, scale_modifier=1.0, background_depth=1.0, background=fake, vertex=fake,
             shs=fake, feature=None, opacity=fake, model=None)
    g.update(geom_over)
    geom = _lib.Geometry(**g)
    flags = _lib.Flags(0, 1, 0, 0, 1, 0, 0)
    return cam, geom, flags, fake


def test_argument_errors_are_reported_before_any_device_work():
    """No GPU here: every call below must fail in validation (negative TS2D_E_* code), not in the CUDA runtime."""
    import ctypes as C

    lib = _lib.load()
    R = C.c_int64(0)
    cam, geom, flags, fake = _dummy_call_structs()
    # reference-shaped errors (extension_interface.cu:53-76)
    for over, code in ((dict(C=4), -4), (dict(gamma=-1.0), -6), (dict(shs=None), -3), (dict(sh_degree=2), -9), (dict(vertex=None), -7),
                       (dict(use_shs=0, feature=None), -2)):
        cam, geom, flags, fake = _dummy_call_structs(**over)
        assert lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, 1 << 30, C.byref(R), None) == code, over
    # shard / primitive
    cam, geom, flags, fake = _dummy_call_structs()
    flags.shard_rank, flags.shard_world = 2, 2
    assert lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, 1 << 30, C.byref(R), None) == -10
    flags.shard_rank, flags.primitive = 0, 7
    assert lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, 1 << 30, C.byref(R), None) == -12
    # parameter-space inputs: f_dc / logits required, f_rest required for M > 1, ratio > 0, SH mode only
    for mi, M, use_shs in ((_lib.ModelInputs(None, None, fake, -1.0, 1.0, 1), 1, 1), (_lib.ModelInputs(fake, None, fake, -1.0, 1.0, 1), 4, 1),
                           (_lib.ModelInputs(fake, fake, None, -1.0, 1.0, 1), 4, 1), (_lib.ModelInputs(fake, fake, fake, -1.0, 0.0, 1), 4, 1),
                           (_lib.ModelInputs(fake, fake, fake, -1.0, 1.0, 1), 4, 0)):
        cam, geom, flags, fake = _dummy_call_structs(M=M, use_shs=use_shs, shs=None, opacity=None, feature=fake if not use_shs else None,
                                                     model=C.cast(C.pointer(mi), C.c_void_p))
        assert lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, 1 << 30, C.byref(R), None) == -13
    assert b"model inputs" in lib.ts2d_error_string(-13)
    # exchange kernels: 2 <= world <= 8, 0 <= rank < world, known operation, aligned base / first / count -- checked before any launch
    for args in ((fake, fake, 7, 64, 64, 0, 1), (fake, fake, 7, 64, 64, 2, 2), (fake, fake, 7, 64, 64, 0, 9), (fake, C.c_void_p(0x1004), 7, 64, 64, 0, 2)):
        assert lib.ts2d_exchange_tiles(*args, None) == -14, args
    assert lib.ts2d_exchange_tiles(None, fake, 7, 64, 64, 0, 2, None) == -7
    for args in ((fake, 2, 8, 0), (fake, 0, 6, 0), (fake, 0, 8, 5), (C.c_void_p(0x1004), 0, 8, 0), (fake, -4, 8, 1)):
        assert lib.ts2d_exchange_allreduce(*args, None) == -14, args
    assert lib.ts2d_exchange_allreduce(fake, 0, 0, 0, None) == 0  # empty slice: nothing to do
    loss = _lib.LossIn(fake, fake, fake)
    assert b"ts2d_exchange" in lib.ts2d_error_string(-14)
    # state blobs too small for what they must hold
    cam, geom, flags, fake = _dummy_call_structs()
    out = _lib.ForwardOut(fake, fake, fake, fake, fake, fake)
    assert lib.ts2d_forward_render(C.byref(cam), C.byref(geom), C.byref(flags), 10**6, fake, fake, 4096, fake, 1 << 30, C.byref(out), None, None) == -8
    assert lib.ts2d_forward(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, 16, fake, 1 << 30, fake, 1 << 30, C.byref(out), None, None) == -8
    bout = _lib.BackwardOut(fake, fake, fake, fake, fake, None)
    assert lib.ts2d_backward(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, fake, 1 << 20, fake, C.byref(loss), C.byref(bout), fake, 64, None) == -8
    # model gradients must come with model inputs and vice versa
    mg = _lib.ModelGrads(fake, None, *([None] * 8), 1)
    bout = _lib.BackwardOut(fake, fake, fake, fake, fake, C.cast(C.pointer(mg), C.c_void_p))
    cam, geom, flags, fake = _dummy_call_structs()
    assert lib.ts2d_backward(C.byref(cam), C.byref(geom), C.byref(flags), fake, fake, fake, 1 << 20, fake, C.byref(loss), C.byref(bout), fake, 1 << 30, None) == -13
    # small utilities
    assert lib.ts2d_downsample(None, fake, 1, 8, 8, 2, None) == -7 and lib.ts2d_downsample(fake, fake, 0, 8, 8, 2, None) == -11
    assert lib.ts2d_downsample_bwd(fake, fake, 1, 8, 8, 0, None) == -11
    # P == 0 short-circuits (extension_interface.cu:130)
    cam, geom, flags, fake = _dummy_call_structs(P=0)
    assert lib.ts2d_forward_geometry(C.byref(cam), C.byref(geom), C.byref(flags), None, None, 0, C.byref(R), None) == 0 and R.value == 0
