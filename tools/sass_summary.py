#!/usr/bin/env python3
"""Instruction-class counts per kernel from the SASS of the built library (static counts: what the code is made of, not what runs):
    python tools/sass_summary.py [path/to/libts2d.so] > profiles/sass_summary_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "triangle_splatting_b200", "lib", "libts2d.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout

CLASSES = [
    ("fp32 fma/mul/add", r"^(FFMA|FMUL|FADD)$"),
    ("fp32 packed x2", r"^(FFMA2|FMUL2|FADD2)$"),
    ("fp32 min/max/cmp/sel", r"^(FMNMX\d*|FSETP|FSET|FSEL|FCHK)$"),
    ("mufu (rcp/ex2/lg2/sqrt)", r"^MUFU$"),
    ("fp64", r"^D(FMA|MUL|ADD|SETP)$"),
    ("int / logic / shift", r"^(IADD3?|IMAD|IMNMX|ISETP|LOP3|SHF|LEA|SEL|POPC|FLO|BREV|PRMT|IABS|VIADD|VIMNMX\d*|I2FP?|F2I|I2I|F2FP?|F2F|PLOP3|P2R|R2P|BMSK|SGXT|IDP|LOP)$"),
    ("shared ld", r"^LDS$"), ("shared st", r"^STS$"), ("shared atom", r"^ATOMS$"),
    ("global ld", r"^(LDG|LD)$"), ("global st", r"^(STG|ST)$"), ("global red/atom", r"^(RED|ATOMG|ATOM)$"),
    ("local ld/st (spill)", r"^(LDL|STL)$"), ("const ld", r"^LDC$"),
    ("multimem", r"^(MULTIMEM|REDG|LDGMC|STGMC)"),
    ("shfl", r"^SHFL$"), ("vote / match / redux", r"^(VOTE|VOTEU|MATCH|REDUX)$"),
    ("barrier / sync", r"^(BAR|WARPSYNC|BSYNC|BSSY|MEMBAR|ERRBAR|DEPBAR|NANOSLEEP|CCTL|FENCE)$"),
    ("branch / call / exit", r"^(BRA|BRX|JMP|CALL|RET|EXIT|BREAK|BPT|KILL)$"),
    ("mov / s2r / cs2r", r"^(MOV|S2R|CS2R|S2UR|R2UR|UMOV|MOVM)$"),
    ("uniform datapath", r"^U[A-Z0-9]+$"),
]
CLASSES = [(n, re.compile(p)) for n, p in CLASSES]

kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "").replace("void ", "")) for o in out]


names = list(kernels)
short = demangle(names)
print(f"# {os.path.relpath(lib, ROOT)}: static SASS instruction classes per kernel (cuobjdump -sass, sm_100a)")
print("# unclassified mnemonics are listed per kernel so nothing hides in 'other'\n")
for name, sh in zip(names, short):
    c = kernels[name]
    total = sum(c.values())
    if total < 16:
        continue
    print(f"{sh}   [{total} instructions]")
    seen = set()
    for cname, rx in CLASSES:
        n = sum(v for k, v in c.items() if rx.match(k))
        seen |= {k for k in c if rx.match(k)}
        if n:
            print(f"    {cname:28s} {n:6d}  {100.0 * n / total:5.1f} %")
    other = {k: v for k, v in c.items() if k not in seen and k != "NOP"}
    if other:
        print("    other: " + ", ".join(f"{k} {v}" for k, v in sorted(other.items(), key=lambda kv: -kv[1])))
    print()
