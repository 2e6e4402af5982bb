#!/usr/bin/env python3
"""Debug aid (GPU): audit one triangle of a config -- which (pixel, triangle) pairs the REFERENCE arithmetic blends (fp32 numpy
restatement on the GPU's own state), per tile / sub-tile, against the coverage bits our forward pass leaves in the instance keys.
    python tools/debug_triangle.py C3 641034 295324"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from triangle_splatting_b200 import _C  # noqa: E402
from triangle_splatting_b200.scenes import make_config  # noqa: E402

dev = torch.device("cuda:0")
sc = make_config(sys.argv[1])
ids = [int(a) for a in sys.argv[2:]]
s = sc.to(dev)
_C.set_exact(False)
fwd = _C.rasterize_triangles(*harness._fwd_args(s))
st = harness.decode_state(s, fwd, dev)
st["radii"] = fwd[2].cpu().numpy()
W, H = sc.cam["image_width"], sc.cam["image_height"]
keys = harness.sorted_instance_keys(fwd, W, H)  # after the forward pass: emit mask minus the bits K7 cleared
gx = (W + 15) // 16
f32 = np.float32
for tid in ids:
    v = st["v2d"][tid].astype(f32)
    area2, op = f32(st["area2"][tid]), f32(sc.opacity[tid, 0].item())
    print(f"tri {tid}: v2d {v.ravel()} area2 {area2} op {op} radii {st['radii'][tid]} rect {st['rect_min'][tid]}..{st['rect_max'][tid]}")
    (x0, y0), (x1, y1) = st["rect_min"][tid], st["rect_max"][tid]
    for ty in range(y0, y1):
        for tx in range(x0, x1):
            tile = ty * gx + tx
            a, b = st["ranges"][tile]
            pos = a + np.nonzero(st["point_list"][a:b] == tid)[0]
            if pos.size != 1:
                print(f"  tile {tile}: instance not found ({pos.size})")
                continue
            pos = int(pos[0])
            rel = pos - a
            mask = int(keys[pos] & 0xFF)
            ys, xs = np.mgrid[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16]
            inside = (xs < W) & (ys < H)
            px, py = xs.astype(f32), ys.astype(f32)
            pv = [(v[k, 0] - px, v[k, 1] - py) for k in range(3)]
            c1 = (pv[1][0].astype(np.float64) * pv[2][1] - (pv[1][1] * pv[2][0]).astype(np.float64)).astype(f32)  # fma(a, b, -fl(c d))
            c2 = (pv[2][0].astype(np.float64) * pv[0][1] - (pv[2][1] * pv[0][0]).astype(np.float64)).astype(f32)
            a1, a2 = c1 / area2, c2 / area2
            a3 = (f32(1) - a1) - a2
            ecc = (np.minimum(np.minimum(a1, a2), a3).astype(np.float64) * -3.0 + 1.0).astype(f32)
            alpha = np.minimum(f32(0.99), op * np.exp(f32(-0.5) * ecc * ecc).astype(f32))
            blends = inside & (ecc >= 0) & (ecc <= 10) & (alpha >= f32(1.0 / 255.0))
            ncon = np.zeros((16, 16), np.int64)
            ncon[inside] = st["n_contrib"][ys[inside], xs[inside]]
            visited = blends & (ncon > rel)
            per_sub = [int(visited[(w >> 1) * 4:(w >> 1) * 4 + 4, (w & 1) * 8:(w & 1) * 8 + 8].sum()) for w in range(8)]
            need = sum(1 << w for w in range(8) if per_sub[w])
            flag = "" if (need & ~mask) == 0 else "   <-- reference blends pairs in a sub-tile whose bit is CLEAR"
            print(f"  tile {tile} pos {pos} rel {rel} (list {b - a}) mask {mask:08b} needed {need:08b} pairs/sub-tile {per_sub} max alpha {alpha[inside].max():.4f}{flag}")
