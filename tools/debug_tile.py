#!/usr/bin/env python3
"""Debug aid (GPU): where do the fast and the mirror backward disagree for ONE triangle?  Renders single tiles (shard = (tile, n_tiles):
only that tile is owned) and then single pixels of the worst tile (upstream gradient zero elsewhere) in both arithmetic modes.
    python tools/debug_tile.py C3 641034"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from triangle_splatting_b200 import _C, distributed  # noqa: E402
from triangle_splatting_b200.scenes import make_config  # noqa: E402

distributed.reduce_accumulators = lambda acc: None  # single process: a "shard" here is just a way to render one tile
dev = torch.device("cuda:0")
sc = make_config(sys.argv[1])
tid = int(sys.argv[2])
s = sc.to(dev)
W, H = sc.cam["image_width"], sc.cam["image_height"]
gx, gy = (W + 15) // 16, (H + 15) // 16


def grads(exact, shard, g_feat):
    _C.set_exact(exact)
    fwd = _C.rasterize_triangles(*harness._fwd_args(s), shard=shard)
    args = list(harness._bwd_args(s, fwd, dev))
    args[19] = g_feat
    bwd = _C.rasterize_triangles_backward(*args, shard=shard)
    _C.set_exact(False)
    return bwd[0][tid].double().cpu().numpy().ravel(), bwd[1][tid].double().cpu().numpy().ravel(), fwd


g_full = s.grads["dL_dout_feature"]
fwd = _C.rasterize_triangles(*harness._fwd_args(s))
st = harness.decode_state(s, fwd, dev)
(x0, y0), (x1, y1) = st["rect_min"][tid], st["rect_max"][tid]
print(f"tri {tid}: rect {x0},{y0}..{x1},{y1}  v2d {st['v2d'][tid].ravel()}")
worst, worst_tile = 0.0, None
for ty in range(y0, y1):
    for tx in range(x0, x1):
        tile = ty * gx + tx
        a_v, a_c, _ = grads(True, (tile, gx * gy), g_full)
        b_v, b_c, _ = grads(False, (tile, gx * gy), g_full)
        d = np.abs(a_v - b_v).max()
        print(f"  tile {tile}: exact dL_dvertex {a_v}\n            fast  dL_dvertex {b_v}\n            max |diff| {d:.3e}   center2D exact {a_c} fast {b_c}")
        if d > worst:
            worst, worst_tile = d, (tx, ty)
tx, ty = worst_tile
tile = ty * gx + tx
print(f"worst tile {tile}: per-pixel audit")
for ly in range(16):
    for lx in range(16):
        px, py = tx * 16 + lx, ty * 16 + ly
        if px >= W or py >= H:
            continue
        g1 = torch.zeros_like(g_full)
        g1[:, py, px] = g_full[:, py, px]
        a_v, _, _ = grads(True, (tile, gx * gy), g1)
        b_v, _, _ = grads(False, (tile, gx * gy), g1)
        d = np.abs(a_v - b_v).max()
        if d > 1e-3 * max(np.abs(a_v).max(), 1e-12) and np.abs(a_v).max() > 0:
            w = (ly >> 2) * 2 + (lx >> 3)
            print(f"   pixel ({px},{py}) sub-tile {w}: exact {a_v[:6]} fast {b_v[:6]} diff {d:.3e}")
