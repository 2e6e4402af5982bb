#!/usr/bin/env python3
"""Generates tests/golden/depth_normal_*.npz from the REFERENCE's own DepthNormalLoss (src/diff_recon/trainers/trainer_utils.py:203-257),
imported from /root/reference in the build container (CPU, fp32 -- the class only runs in fp32: its Scharr kernels are fp32 tensors).
The module imports torchmetrics and simple_knn at the top, neither of which the loss uses: both are stubbed.  Run once; the fixtures
travel, the reference does not."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/diff_recon/trainers/trainer_utils.py"


def load_reference_module(path=REF):
    for name in ("torchmetrics", "torchmetrics.image", "torchmetrics.image.lpip", "simple_knn"):
        sys.modules.setdefault(name, types.ModuleType(name))

    class _Unused:
        def __init__(self, *a, **k):
            pass

        def to(self, *a, **k):
            return self

    sys.modules["torchmetrics.image.lpip"].LearnedPerceptualImagePatchSimilarity = _Unused
    sys.modules["simple_knn"].nearestNeighbor = None
    spec = importlib.util.spec_from_file_location("ref_trainer_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scene(h, w, seed):
    """A depth map with smooth slopes, a few steps (large Scharr gradients: the masked tail) and noise; a perturbed normal map."""
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    depth = 4.0 + 0.02 * x + 0.03 * y + 0.5 * torch.sin(x / 7.0) * torch.cos(y / 5.0) + 0.02 * torch.rand(h, w, generator=g)
    depth = depth + 1.5 * (x > 0.6 * w).float() - 0.8 * (y > 0.7 * h).float()
    normal = torch.randn(3, h, w, generator=g) * 0.3 + torch.tensor([0.1, -0.2, -1.0]).view(3, 1, 1)
    return depth, normal


CASES = {"half_38x50": (38, 50, 0.5), "half_odd_37x53": (37, 53, 0.5), "full_21x25": (21, 25, None), "half_96x128": (96, 128, 0.5)}

if __name__ == "__main__":
    ref = load_reference_module()
    for name, (h, w, sf) in CASES.items():
        depth, normal = scene(h, w, seed=h * 1000 + w)
        d, n = depth.clone().requires_grad_(True), normal.clone().requires_grad_(True)
        loss = ref.DepthNormalLoss(scale_factor=sf)(d, n, 0.55, 0.41)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, f"depth_normal_{name}.npz"), depth=depth.numpy(), normal=normal.numpy(), tan_fovx=0.55, tan_fovy=0.41,
                            scale_factor=np.float64(-1.0 if sf is None else sf), loss=np.float32(loss.item()), g_depth=d.grad.numpy(),
                            g_normal=n.grad.numpy())
        print(name, float(loss.detach()))
