#!/usr/bin/env python3
"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference CUDA
extension (oracle/_ref, built by oracle/build_ref.py from /root/reference) on a B200.

    gpurun -- 'python tests/golden/make_golden.py [2D] [3D]'   # writes gpurun_out/golden/*.npz (3D files are 3d_<name>.npz)
    cp gpurun_out/golden/*.npz tests/golden/              # then commit

The reference ships no fixtures of its own; these files are what pins the CPU
oracle (tests/test_oracle_golden.py) and, bit-exactly for the integer fields, our CUDA path
(tests/test_gpu_parity.py).  Inputs are regenerated from the seeds in tests/harness.GOLDEN_SCENES,
so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    which = [a for a in sys.argv[1:] if a in ("2D", "3D")] or ["2D", "3D"]
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for primitive in which:
        ref = harness.load_reference(primitive)
        if ref is None:
            raise SystemExit(f"oracle/_ref ({primitive}) is not built (run oracle/build_ref.py where /root/reference exists)")
        scenes = harness.GOLDEN_SCENES if primitive == "2D" else harness.GOLDEN_SCENES_3D
        for name in scenes:
            sc = harness.golden_scene(name, primitive)
            out = harness.run_reference(sc, dev, backward=True, ref=ref, primitive=primitive)
            # inputs checksum so a consumer can verify it regenerated the same scene
            out["input_checksum"] = np.float64(sc.vertex.double().sum().item() + sc.opacity.double().sum().item())
            path = os.path.join(outdir, os.path.basename(harness.golden_path(name, primitive)))
            np.savez_compressed(path, **out)
            print(primitive, name, "P", sc.P, "R", int(out["num_rendered"]), "visible", int((out["radii"] > 0).sum()),
                  "img_mean", float(out["out_feature"].mean()), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
