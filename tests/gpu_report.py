#!/usr/bin/env python3
"""GPU-side parity report (not a test): quantifies, per output tensor, how far apart are
  ref-A vs ref-B   two runs of the unmodified reference (its gradient sums use fp32 atomics: run-to-run noise)
  exact vs ref     our mirror-arithmetic kernels vs the reference
  fast  vs ref     our fast kernels vs the reference
  *     vs f64     each of them vs the fp64 CPU oracle ("truth"), small scenes only
with the acceptance metric of SURVEY.md section 8(d): max |a-b| / max(|b|, 1e-3 RMS(b)), plus the fraction of
entries above 1e-5.  Output: gpurun_out/report.txt (copied to profiles/ when committed)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from harness import frac_above, mismatch_count, rel_err  # noqa: E402
from triangle_splatting_b200 import _C  # noqa: E402
from triangle_splatting_b200.scenes import make_scene  # noqa: E402

KEYS = ("out_feature", "depth", "normal", "contrib_sum", "contrib_max", "final_T", "dL_dvertex", "dL_dcenter2D", "dL_dshs", "dL_dfeature",
        "dL_dopacity")


def main():
    dev = torch.device("cuda:0")
    prim = "3D" if "3D" in sys.argv[1:] else "2D"  # `python tests/gpu_report.py 3D` -> gpurun_out/report_3d.txt
    ref = harness.load_reference(prim)
    out = open(os.path.join(ROOT, "gpurun_out", "report.txt" if prim == "2D" else "report_3d.txt"), "w")

    def emit(*a):
        s = " ".join(str(x) for x in a)
        print(s)
        out.write(s + "\n")

    scenes = {n: harness.golden_scene(n, prim) for n in (harness.GOLDEN_SCENES if prim == "2D" else harness.GOLDEN_SCENES_3D)}
    if prim == "3D":
        scenes.pop("sh0_plain")  # rich_info=False: the reference's 3D backward crashes (harness.run_reference)
    scenes["mid_sh3_rich_200k_800x600"] = make_scene("mid", 200_000, 800, 600, sh_degree=3, rich_info=True, geometry_grads=True, seed=21)
    scenes["gamma7_ste_50k_640x480"] = make_scene("g7", 50_000, 640, 480, sh_degree=0, rich_info=True, gamma=7.0, opacity_ste=0.3, seed=22)
    for name, sc in scenes.items():
        runs = {}
        if ref is not None:
            runs["refA"] = harness.run_reference(sc, dev, ref=ref, primitive=prim)
            runs["refB"] = harness.run_reference(sc, dev, ref=ref, primitive=prim)
        old = _C.set_exact(True)
        runs["exact"] = harness.run_ours(sc, dev, primitive=prim)
        _C.set_exact(False)
        runs["fast"] = harness.run_ours(sc, dev, primitive=prim)
        _C.set_exact(old)
        if sc.P <= 200000:
            runs["f64"] = harness.run_oracle(sc, "f64", primitive=prim)
        base = runs.get("refA", runs["exact"])
        emit(f"== {name}: P={sc.P} R={int(base['num_rendered'])} visible={int((base['radii'] > 0).sum())} rich={sc.rich_info} gamma={sc.gamma}")
        for who in ("exact", "fast"):
            ints = {k: mismatch_count(runs[who][k], base[k]) for k in harness.INT_KEYS if k in base and k in runs[who]}
            bits = {k: mismatch_count(np.asarray(runs[who][k]).view(np.uint32), np.asarray(base[k]).view(np.uint32))
                    for k in ("out_feature", "final_T", "depth", "normal") if k in base and k in runs[who]}
            emit(f"   {who:5s} integer mismatches vs ref: {ints}")
            emit(f"   {who:5s} forward outputs differing in bits vs ref: {bits}")
        pairs = [("refB", "refA"), ("exact", "refA"), ("fast", "refA"), ("fast", "exact")]
        if "f64" in runs:
            pairs += [("refA", "f64"), ("exact", "f64"), ("fast", "f64")]
        emit("   %-13s " % "tensor" + " ".join("%-21s" % f"{a}~{b}" for a, b in pairs))
        for k in KEYS:
            row = []
            for a, b in pairs:
                if a in runs and b in runs and k in runs[a] and k in runs[b]:
                    row.append("%.1e (%.1e)       " % (rel_err(runs[a][k], runs[b][k]), frac_above(runs[a][k], runs[b][k], 1e-5, 1e-3)))
                else:
                    row.append("-" + " " * 20)
            emit("   %-13s " % k + " ".join(row))
    out.close()


if __name__ == "__main__":
    main()
