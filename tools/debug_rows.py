#!/usr/bin/env python3
"""Debug aid (GPU): fast vs mirror kernels on the golden scenes, per gradient tensor and per triangle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from triangle_splatting_b200 import _C  # noqa: E402

dev = torch.device("cuda:0")
for prim, scenes in (("2D", harness.GOLDEN_SCENES), ("3D", harness.GOLDEN_SCENES_3D)):
    for name in scenes:
        sc = harness.golden_scene(name, prim)
        _C.set_exact(True)
        a = harness.run_ours(sc, dev, primitive=prim)
        _C.set_exact(False)
        b = harness.run_ours(sc, dev, primitive=prim)
        b2 = harness.run_ours(sc, dev, primitive=prim)
        line = [f"{prim}/{name}: R={int(a['num_rendered'])}"]
        for k in harness.GRAD_KEYS:
            line.append(f"{k} {harness.rel_err(b[k], a[k]):.1e} (rerun {'same' if np.array_equal(b[k], b2[k]) else 'DIFFERS'})")
        print("  ".join(line), flush=True)
        gv_a, gv_b = a["dL_dvertex"].reshape(sc.P, -1), b["dL_dvertex"].reshape(sc.P, -1)
        scale = 1e-3 * np.sqrt(np.mean(gv_a ** 2)) + 1e-30
        err = (np.abs(gv_b - gv_a) / np.maximum(np.abs(gv_a), scale)).max(axis=1)
        bad = np.nonzero(err > 5e-2)[0]
        if bad.size:
            tt = a["tiles_touched"]
            order = np.argsort(a["tri_depth"] + np.where(a["radii"] > 0, 0, 1e30), kind="stable")
            rank = np.empty(sc.P, np.int64)
            rank[order] = np.arange(sc.P)
            print(f"   {bad.size} triangles off by > 5e-2; ids {bad[:12]} ranks {rank[bad[:12]]} tiles {tt[bad[:12]]} of {int((a['radii'] > 0).sum())} visible")
