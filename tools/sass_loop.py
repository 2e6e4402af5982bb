#!/usr/bin/env python3
"""Instruction budget of a hot loop, from the SASS of the built objects (static; no GPU needed):

    python tools/sass_loop.py > profiles/r02/walk_loop_budget.txt

For the two composite kernels (the instantiation the benchmark runs: RICH, gamma == 1, one warp per CTA) it lists the instructions
on the common path of one walk iteration -- one (warp, list entry) visit in which some lane blends and nothing falls into a
decision band -- i.e. the loop body without the out-of-line blocks (exact re-evaluation, saturation handling, panel flush), grouped
by what they do.  Together with the visit counts of the bench line (`work`) this is the compute model of K7 / K8: they are bound by
instruction issue (ncu: 76-78 % of the issue slots), so instructions per visit x visits / issue rate is their time.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "triangle_splatting_b200", "lib", "obj")

GROUPS = [
    ("fp32 add/mul/fma (scalar)", r"^(FADD|FMUL|FFMA)$"),
    ("fp32 packed (2 per slot)", r"^(FADD2|FMUL2|FFMA2)$"),
    ("fp32 min/max/compare/select", r"^(FMNMX3?|FSETP|FSEL)$"),
    ("MUFU (ex2 / rcp)", r"^MUFU$"),
    ("shared-memory load", r"^LDS$"),
    ("shared-memory store", r"^STS$"),
    ("vote / ballot", r"^(VOTE|VOTEU)$"),
    ("integer / predicate / address", r"^(IADD3|IMAD|ISETP|LOP3|PLOP3|LEA|SEL|VIADD|PRMT|MOV|HFMA2|UIADD3|UISETP|UMOV|UIMAD|R2UR)$"),
    ("branch / convergence", r"^(BRA|BSSY|BSYNC|WARPSYNC|NOP)$"),
]


def sass_of(obj, pattern):
    names = subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout
    fn = [m.group(1) for m in re.finditer(r"Function : (\S+)", names) if pattern in m.group(1)]
    assert len(fn) == 1, (pattern, fn)
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn[0], obj], check=True, capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def mnemonic(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0]


COND = re.compile(r"^(?:@!?U?P\d+\s+BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?|BRA(?:\.U)?\s+!?U?P\d+,\s*)0x([0-9a-f]+)")
UNCOND = re.compile(r"^BRA(?:\.U)?\s+0x([0-9a-f]+)")


def trace(ins, start, taken, limit=400):
    """Follow the control flow from `start`: unconditional branches are followed, the n-th conditional branch met (n = 1, 2, ...) is taken
    iff n is in `taken`; stops at the first backward branch that is taken by this rule or after `limit` instructions.  Returns the list
    of (address, text, note) executed -- predicated non-branch instructions count as issued (they occupy an issue slot either way)."""
    index = {a: i for i, (a, _) in enumerate(ins)}
    i, n, path = index[start], 0, []
    while len(path) < limit:
        a, t = ins[i]
        m = COND.match(t)
        if m:
            n += 1
            tgt = int(m.group(1), 16)
            go = n in taken
            path.append((a, t, f"conditional #{n}: " + ("taken" if go else "not taken")))
            if go:
                if tgt <= a:
                    return path
                i = index[tgt]
            else:
                i += 1
            continue
        m = UNCOND.match(t)
        path.append((a, t, ""))
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a:
                return path
            i = index[tgt]
        else:
            i += 1
    return path


def report(title, obj, pattern, head_text, head_nth, taken, notes):
    ins = sass_of(os.path.join(OBJ, obj), pattern)
    heads = [a for a, t in ins if head_text in t]
    start = heads[head_nth]
    path = trace(ins, start, taken)
    print(f"== {title}")
    print(f"   function ...{pattern}..., from 0x{start:04x}: {len(path)} issue slots on the common path")
    counts = {}
    for _, t, _ in path:
        mn = mnemonic(t)
        for g, pat in GROUPS:
            if re.match(pat, mn):
                counts[g] = counts.get(g, 0) + 1
                break
        else:
            counts["other: " + mn] = counts.get("other: " + mn, 0) + 1
    for g, _ in GROUPS:
        if g in counts:
            print(f"     {g:36s} {counts[g]:3d}")
    for g in sorted(k for k in counts if k.startswith("other")):
        print(f"     {g:36s} {counts[g]:3d}")
    for n in notes:
        print("   " + n)
    print("   path:")
    for a, t, note in path:
        print(f"     /*{a:04x}*/ {t:78s}" + (f"  <- {note}" if note else ""))
    print()


if __name__ == "__main__":
    print("# Instruction budget of the walk loops (static, from cuobjdump -sass of triangle_splatting_b200/lib/obj/*.o; tools/sass_loop.py)")
    print("# common path = one (warp, list entry) visit in which a lane blends and no lane is inside a decision band; blocks that contain a CALL,")
    print("# a global load or shuffles (exact re-evaluation, exact-transmittance re-walk, panel flush) are out of line and not counted")
    print()

    # K7: the iteration starts at `if (!done)`: LOP3.LUT P0, RZ, Rdone, 0xff.  Conditional branches on the common path: #1 done -> not taken,
    # #2 "outside every decision band" -> taken, #3 "no hit" -> not taken, #4 "no lane reached the 1e-4 cut" -> taken, #5 loop back -> taken
    report("K7 k_render_fwd_fast<RICH, gamma == 1, 1 warp per CTA>: one walk iteration", "ts2d_render_fwd_fast.cu.o", "k_render_fwd_fastILb1ELb1ELi1E",
           "LOP3.LUT P0, RZ, R8, 0xff", 0, {2, 4, 5},
           ["per C3 frame: ~14 M visits; ncu: 1.010 G warp-instructions per launch, 75.7 % issue-active, 1.216 ms"])
    # K8: the iteration starts with the LDS.128 of the entry's vertex words.  #1 inside a band -> not taken, #2 "live and arg-min tie" -> not taken,
    # (BRA to the hit test), #3 "no hit" -> not taken, #4 "no geometry gradients in this sub-tile" -> taken (C3), #5 "panel not full" -> taken
    # (7 of 8 iterations), #6 loop back -> taken
    report("K8 k_render_bwd_fast<RICH, gamma == 1, 1 warp per CTA>: one walk iteration (phase 1, colour-only sub-tile)", "ts2d_render_bwd_fast.cu.o",
           "k_render_bwd_fastILb1ELb1ELi1E", "LDS.128 R20, [UR", 0, {4, 5, 6},
           ["phase 2 (bwd_flush_panel, out of line, every 8 rows) adds ~31 instructions per row;",
            "per C3 frame: 11.2 M visits (= backward rows); ncu: 1.596 G warp-instructions per launch, 77.9 % issue-active, 1.832 ms"])
    sys.exit(0)
