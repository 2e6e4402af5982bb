"""What DESIGN.md says about the built library, checked mechanically on the SASS of triangle_splatting_b200/lib/libts2d.so (CPU: cuobjdump
only reads the fatbin).  These are static facts about the code that ships -- the dynamic counterparts (two runs bit-identical, parity with the
reference) are GPU tests.

  * the library is sm_100a code and nothing else;
  * no kernel of the gradient path accumulates with global atomics: K8 (fast, 2D and 3D), the row marking, the row reduction and K9 contain
    no RED / REDG / ATOM / ATOMG instruction (the reference: 10-16 atomicAdd per pixel and triangle, R2D/src/backward.cu:412,482-490);
  * no library kernels: no cub / thrust symbol is linked in (the reference sorts and scans with CUB, R2D/src/rasterizer.cu:186,211);
  * the instantiations the benchmark runs (RICH, one warp per CTA) do not spill;
  * every kernel of a frame carries the dependent-launch pair (PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait);
  * the exchange kernels really are multimem code (NVSwitch multicast stores / in-switch reduction).
"""
import re
import shutil
import subprocess
from functools import lru_cache
from pathlib import Path

import pytest

LIB = Path(__file__).resolve().parent.parent / "triangle_splatting_b200" / "lib" / "libts2d.so"
ATOMIC = re.compile(r"^(RED|REDG|ATOM|ATOMG)(\.|$)")  # global reductions / atomics (ATOMS = shared memory, REDUX = warp reduce: not these)

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None, reason="needs cuobjdump and c++filt")


@lru_cache(maxsize=1)
def kernels():
    """demangled kernel name -> list of SASS mnemonics (with their .suffixes)"""
    from triangle_splatting_b200 import build

    build.build(force=False)
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], check=True, capture_output=True, text=True).stdout
    out, cur, archs = {}, None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            archs.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    names = list(out)
    dem = subprocess.run(["c++filt"] + names, check=True, capture_output=True, text=True).stdout.strip().split("\n")
    assert len(dem) == len(names)
    res = {d: out[n] for d, n in zip(dem, names)}
    res["__archs__"] = sorted(archs)
    return res


def matching(pattern):
    ks = {k: v for k, v in kernels().items() if k != "__archs__" and re.search(pattern, k)}
    assert ks, f"no kernel matches {pattern!r}"
    return ks


def test_sm100a_only():
    assert kernels()["__archs__"] == ["sm_100a"], kernels()["__archs__"]


@pytest.mark.parametrize("pattern", [r"k_render_bwd_fast<", r"k_render3d_bwd_fast<", r"k_bwd_rows_mark", r"k_bwd_rows_reduce", r"k_preprocess_bwd<",
                                     r"k_preprocess3d_bwd<"])
def test_gradient_path_has_no_global_atomics(pattern):
    for name, ins in matching(pattern).items():
        bad = sorted({i for i in ins if ATOMIC.match(i)})
        assert not bad, f"{name}: global atomics {bad} in the atomics-free gradient path"


def test_mirror_backward_keeps_the_reference_atomics():
    """Sanity of the check above: the op-for-op mirror of the reference's backward (flags.exact) does use them."""
    for name, ins in matching(r"k_render_bwd<").items():
        assert any(ATOMIC.match(i) and ".F32" in i for i in ins), name


def test_no_library_kernels_linked():
    syms = subprocess.run(["nm", "-C", "--defined-only", str(LIB)], check=True, capture_output=True, text=True).stdout
    assert "cub::" not in syms and "thrust::" not in syms
    for k in kernels():
        assert "cub::" not in k and "thrust::" not in k, k


def test_fast_forward_atomics_are_integer_only():
    """K7 keeps three global reductions: contrib_sum as a 64-bit fixed-point ADD, contrib_max as an unsigned MAX on the float's bit pattern, and the
    AND that clears a sub-tile coverage bit -- integer operations, so their result does not depend on the order they arrive in."""
    for name, ins in {**matching(r"k_render_fwd_fast<"), **matching(r"k_render3d_fwd_fast<")}.items():
        at = sorted({i for i in ins if ATOMIC.match(i)})
        assert at, name
        assert not [i for i in at if ".F32" in i or ".F64" in i or ".F16" in i], f"{name}: floating-point atomics {at}"


@pytest.mark.parametrize("pattern", [r"k_render_fwd_fast<true, (true|false), 1>", r"k_render_bwd_fast<true, (true|false), 1>",
                                     r"k_render3d_fwd_fast<true, (true|false), 1>", r"k_render3d_bwd_fast<true, (true|false), 1>",
                                     r"k_preprocess_bwd<", r"k_emit_warp<", r"k_bwd_rows_reduce", r"k_bwd_rows_mark"])
def test_shipped_instantiations_do_not_spill(pattern):
    """(k_radix_pass, k_radix_hist and k_preprocess index small arrays dynamically -- the eight look-back words of a round, the per-digit
    argument arrays -- which ptxas keeps in local memory: they are not in this list.)"""
    for name, ins in matching(pattern).items():
        spills = [i for i in ins if re.match(r"^(LDL|STL)(\.|$)", i)]
        assert not spills, f"{name}: {len(spills)} local-memory accesses"


def test_every_frame_kernel_has_the_dependent_launch_pair():
    frame = r"k_preprocess|k_radix_|k_scan_|k_emit_warp|k_tile_tables|k_render_(fwd|bwd)_fast|k_render3d_(fwd|bwd)_fast|k_contrib_finish|k_bwd_rows_"
    ks = matching(frame)
    assert len(ks) >= 30
    for name, ins in ks.items():
        assert "PREEXIT" in ins and "ACQBULK" in ins, f"{name}: griddepcontrol pair missing"
        first_mem = next((j for j, i in enumerate(ins) if re.match(r"^(LDG|STG|LD|ST|RED|REDG|ATOM|ATOMG)(\.|$)", i)), len(ins))
        assert ins.index("ACQBULK") < first_mem, f"{name}: a global access is scheduled before griddepcontrol.wait"


def test_exchange_kernels_are_multimem_code():
    """multimem.ld_reduce is LDGMC (in-switch reduction), multimem.st a system-scope STG on the multicast address."""
    for name, ins in matching(r"k_exchange_allreduce<").items():
        assert any(i.startswith("LDGMC") for i in ins), f"{name}: no in-switch reduction load"
        assert any(i.startswith("STG") and i.endswith("STRONG.SYS") for i in ins), name
    for name, ins in matching(r"k_exchange_tiles").items():
        assert any(i.startswith("STG.E.128") and i.endswith("STRONG.SYS") for i in ins), f"{name}: no 16-byte system-scope store"
