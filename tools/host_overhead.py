#!/usr/bin/env python3
"""Host-side cost of one forward + backward through the public API (ctypes over the C ABI), measured where the GPU work is
negligible: a 200-triangle scene on a 64x64 frame.  Wall time per step with a synchronise at the end of every step (= host issue
time + launch latencies of the ~25 kernels of a frame) and the CPU time spent in the calls themselves (no synchronise inside the loop)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from triangle_splatting_b200 import TriangleRasterizationSettings, TriangleRasterizer  # noqa: E402
from triangle_splatting_b200.scenes import make_scene  # noqa: E402

dev = torch.device("cuda:0")
sc = make_scene("tiny", 200, 64, 64, sh_degree=3, rich_info=True, geometry_grads=True, seed=3)
s = sc.to(dev)
vertex = s.vertex.clone().requires_grad_(True)
shs = s.shs.clone().requires_grad_(True)
opacity = s.opacity.clone().requires_grad_(True)
c2d = torch.zeros((s.P, 2), device=dev, requires_grad=True)
rast = TriangleRasterizer(TriangleRasterizationSettings(**s.settings_kwargs()))


def step():
    out = rast.forward(vertex=vertex, center2D=c2d, opacity=opacity, shs=shs)
    loss = (out[0] * s.grads["dL_dout_feature"]).sum()
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    step()
    torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(n):
    step()
t2 = time.perf_counter()
torch.cuda.synchronize()
print(f"forward + backward, 200 triangles, 64x64: {1e6 * (t1 - t0) / n:.0f} us per step with a synchronise per step, "
      f"{1e6 * (t2 - t1) / n:.0f} us of host time per step when the queue is left to run")
from triangle_splatting_b200 import _C  # noqa: E402

nat = _C.native()
print(f"host layer: {'compiled _C_native' if nat is not None else 'ctypes'}; counter handles created: "
      f"{nat.counters_created() if nat is not None else _C.FrameCounters.created} over {20 + 2 * n} steps")
