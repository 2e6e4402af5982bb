#!/usr/bin/env python3
"""Debug aid (GPU): the worst dL_dcenter2D / dL_dvertex entries of ours vs the truth at C3, with the triangle's geometry."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from triangle_splatting_b200.scenes import make_config  # noqa: E402

dev = torch.device("cuda:0")
sc = make_config(sys.argv[1] if len(sys.argv) > 1 else "C3")
ref = harness.run_reference(sc, dev, ref=harness.load_reference("2D"))
ours = harness.run_ours(sc, dev)
truth = harness.run_truth(sc, ref)
for key in ("dL_dcenter2D", "dL_dvertex"):
    t = np.asarray(truth[key], np.float64).reshape(sc.P, -1)
    eps = 1e-3 * np.sqrt(np.mean(t * t))
    eo = np.abs(ours[key].reshape(sc.P, -1) - t) / np.maximum(np.abs(t), eps)
    er = np.abs(ref[key].reshape(sc.P, -1) - t) / np.maximum(np.abs(t), eps)
    print(f"{key}: rms {np.sqrt(np.mean(t * t)):.3e} floor {eps:.3e}")
    for idx in np.argsort(eo.max(axis=1))[::-1][:6]:
        v = ref["v2d"][idx]
        print(f"  tri {idx}: err ours {eo[idx].max():.2e} ref {er[idx].max():.2e} | truth {t[idx]} ours {ours[key].reshape(sc.P, -1)[idx]} ref {ref[key].reshape(sc.P, -1)[idx]}")
        print(f"     area2 {ref['area2'][idx]:.4e} opacity {float(sc.opacity[idx]):.4f} radii {ref['radii'][idx]} tiles {ref['tiles_touched'][idx]} v2d {v.ravel()} depth {ref['tri_depth'][idx]:.3f}")
        print(f"     dL_dvertex truth |max| {np.abs(truth['dL_dvertex'][idx]).max():.3e} center2D truth {truth['dL_dcenter2D'][idx]}")
