"""Static check of the launch chain (CPU): every kernel that libts2d launches through ts2d_launch / ts2d_launch_chained may be
scheduled while its predecessor in the stream is still running (programmatic dependent launch, csrc/ts2d_common.cuh), so each of
them has to execute ts2d_grid_chain() -- griddepcontrol.launch_dependents + griddepcontrol.wait -- BEFORE its first access to
global memory.  A kernel that forgot the wait would race with its predecessor only on the GPU and only sometimes; this test reads
the sources instead."""
import re
from pathlib import Path

CSRC = Path(__file__).resolve().parent.parent / "triangle_splatting_b200" / "csrc"
SOURCES = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")))


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _launched_kernels():
    names = {}
    for p in SOURCES:
        t = _strip_comments(p.read_text())
        for m in re.finditer(r"\bts2d_launch(_chained)?\(\s*([A-Za-z_]\w*)", t):
            if m.group(2) in ("kernel", "void"):  # the launchers' own definitions
                continue
            names.setdefault(m.group(2), set()).add("chained" if m.group(1) else "plain")
    return names


def _kernel_bodies(name: str):
    """(file, parameter text, body text) of every __global__ definition called `name` (templates have one body)."""
    out = []
    for p in SOURCES:
        t = _strip_comments(p.read_text())
        for m in re.finditer(r"__global__[^;{]*?\b" + re.escape(name) + r"\s*\(", t):
            i = m.end()
            depth, j = 1, i
            while depth:  # closing parenthesis of the parameter list
                depth += {"(": 1, ")": -1}.get(t[j], 0)
                j += 1
            params = t[i:j - 1]
            k = t.index("{", j)
            depth, e = 1, k + 1
            while depth:
                depth += {"{": 1, "}": -1}.get(t[e], 0)
                e += 1
            out.append((p.name, params, t[k + 1:e - 1]))
    return out


def _first_global_access(params: str, body: str) -> int:
    """Offset in `body` of the first thing that can touch global memory (conservative, textual)."""
    ptrs = re.findall(r"\*\s*(?:__restrict__\s*)?(\w+)\s*(?:,|$)", params)
    structs = re.findall(r"\b(?:RadixHistArgs|RadixPassArgs|Load\w*|Load)\s+(\w+)\s*(?:,|$)", params)
    pats = [r"__ldg\s*\(", r"\batomic\w*\s*\(", r"\brs_count\s*\(", r"\bld_relaxed_u64\s*\(", r"\bst_relaxed_u64\s*\(", r"\bldg_stream128\s*\(",
            r"\bst_stream128\s*\("]
    for q in ptrs:
        pats += [r"\b" + q + r"\s*\[", r"\*\s*" + q + r"\b", r"\b" + q + r"\s*\+"]
    for q in structs:
        pats += [r"\b" + q + r"\s*\(", r"\b" + q + r"\.\w+\s*\["]
    first = len(body)
    for pat in pats:
        m = re.search(pat, body)
        if m:
            first = min(first, m.start())
    return first


def test_chain_kernels_wait_before_touching_global_memory():
    kernels = _launched_kernels()
    assert len(kernels) >= 15, f"expected the whole frame's kernels behind ts2d_launch*, found {sorted(kernels)}"
    assert any("chained" in v for v in kernels.values())
    checked = 0
    for name in sorted(kernels):
        bodies = _kernel_bodies(name)
        assert bodies, f"no __global__ definition found for launched kernel {name}"
        for fname, params, body in bodies:
            pos = body.find("ts2d_grid_chain()")
            assert pos >= 0, f"{fname}: {name} is launched through ts2d_launch* but never calls ts2d_grid_chain()"
            first = _first_global_access(params, body)
            assert pos < first, f"{fname}: {name} touches global memory before ts2d_grid_chain(): ...{body[max(0, first - 60):first + 40]!r}"
            checked += 1
    assert checked >= len(kernels)


def test_only_short_kernels_are_chained():
    """Measured policy (profiles/r02/pdl_ab): the attribute only inside the sort / scan / table chains."""
    kernels = _launched_kernels()
    chained = {k for k, v in kernels.items() if "chained" in v}
    assert chained == {"k_radix_pass", "k_scan_sums", "k_scan_block_sums", "k_scan_apply", "k_tile_tables"}, chained
    for k in chained:
        assert kernels[k] == {"chained"}, f"{k} is launched both ways"
