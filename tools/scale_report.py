#!/usr/bin/env python3
"""Parity + work statistics at the BENCHMARKED sizes (not a test; run on a B200 box):

    python tools/scale_report.py [C2 C3 C4_3D C5_3D ...]   ->  gpurun_out/scale_report.txt / .json

Per scene: the unmodified reference twice (refA, refB: its gradient sums use fp32 atomics, so this is its run-to-run
spread), our fast kernels twice (fastA, fastB: run-to-run of ours), our mirror kernels (exact).  Reported: integer
mismatches vs the reference, the SURVEY 8(d) metric max |a-b| / max(|b|, 1e-3 RMS(b)) and its upper percentiles for every
float output and gradient, and the work statistics the bench line quotes (list lengths, sum of n_contrib, sub-tile visits).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
from harness import mismatch_count  # noqa: E402
from triangle_splatting_b200 import _C  # noqa: E402
from triangle_splatting_b200.scenes import make_config  # noqa: E402

FLOAT_KEYS = ("out_feature", "depth", "normal", "contrib_sum", "contrib_max", "final_T", "dL_dvertex", "dL_dcenter2D", "dL_dshs", "dL_dfeature",
              "dL_dopacity")

SCENES = {
    "C2": ("2D", dict()),
    "C3": ("2D", dict()),
    "C2_geo": ("2D", dict(geometry_grads=True)),
    "C4_3D": ("3D", dict()),
    "C5_3D": ("3D", dict()),
    "C5_2D": ("2D", dict()),
}


def err_stats(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    if b.size == 0:
        return None
    eps = 1e-3 * float(np.sqrt(np.mean(b * b))) + 1e-30
    e = np.abs(a - b) / np.maximum(np.abs(b), eps)
    q = np.quantile(e, [0.5, 0.99, 0.9999]) if e.size > 1 else [e[0]] * 3
    return dict(max=float(e.max()), p9999=float(q[2]), p99=float(q[1]), med=float(q[0]), above_1e5=float(np.mean(e > 1e-5)),
                above_1e4=float(np.mean(e > 1e-4)))


def work_stats(run, fwd_state, W, H):
    """list lengths, n_contrib sums, sub-tile (8x4) visits from the coverage masks carried in our sorted instance keys."""
    rng = run["ranges"].astype(np.int64)
    lens = rng[:, 1] - rng[:, 0]
    ncon = run["n_contrib"].astype(np.int64)
    out = dict(R=int(run["num_rendered"]), visible=int((run["radii"] > 0).sum()), tiles=int(lens.size), list_len_mean=float(lens.mean()),
               list_len_max=int(lens.max()), pairs_sum_n_contrib=int(ncon.sum()), n_contrib_mean=float(ncon.mean()), pairs_upper_256R=int(256 * int(run["num_rendered"])))
    if fwd_state is not None:
        R = out["R"]
        tkey1 = harness.sorted_instance_keys(fwd_state, W, H)
        masks = tkey1 & 0xFF
        bits = np.unpackbits(masks.astype(np.uint8)[:, None], axis=1, bitorder="little")  # [R, 8]
        out["subtile_visits_mask_bits"] = int(bits.sum())
        # truncated at each sub-tile's own last contributor (what the backward walks)
        gx, gy = (W + 15) // 16, (H + 15) // 16
        nc = np.zeros((gy * 16, gx * 16), np.int64)
        nc[:H, :W] = ncon
        sub = nc.reshape(gy, 4, 4, gx, 2, 8).max(axis=(2, 5))  # [gy, sy(4), gx, sx(2)]
        lastw = sub.transpose(0, 2, 1, 3).reshape(gy * gx, 8)  # w = sy * 2 + sx
        tile = (tkey1 >> 8).astype(np.int64)
        rel = np.arange(R, dtype=np.int64) - rng[tile, 0]
        live = bits.astype(bool) & (rel[:, None] < lastw[tile])
        out["bwd_rows"] = int(live.sum())
        out["bwd_rows_per_instance"] = float(live.sum() / max(R, 1))
    return out


def main():
    names = [a for a in sys.argv[1:] if a in SCENES] or ["C2", "C3"]
    dev = torch.device("cuda:0")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    txt = open(os.path.join(ROOT, "gpurun_out", "scale_report.txt"), "a")
    allres = {}

    def emit(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        txt.write(s + "\n")
        txt.flush()

    for name in names:
        prim, kw = SCENES[name]
        t0 = time.time()
        sc = make_config(name.split("_")[0], **kw)
        ref = harness.load_reference(prim)
        runs = {}
        if ref is not None:
            runs["refA"] = harness.run_reference(sc, dev, ref=ref, primitive=prim)
            runs["refB"] = harness.run_reference(sc, dev, ref=ref, primitive=prim)
        old = _C.set_exact(False)
        s = sc.to(dev)
        fwd = _C.rasterize_triangles(*harness._fwd_args(s), primitive=prim)
        torch.cuda.synchronize()
        runs["fastA"] = harness.run_ours(sc, dev, primitive=prim)
        runs["fastB"] = harness.run_ours(sc, dev, primitive=prim)
        _C.set_exact(True)
        runs["exact"] = harness.run_ours(sc, dev, primitive=prim)
        _C.set_exact(old)
        base = runs.get("refA", runs["exact"])
        if "--truth" in sys.argv:  # fp64 values, the reference's fp32 decisions (oracle kind f64d): comparable entry by entry at any size
            t1 = time.time()
            runs["truth"] = harness.run_truth(sc, base, prim)
            emit(f"   [truth: {time.time() - t1:.0f}s on {os.cpu_count()} host threads]")
        ws = work_stats(runs["fastA"], fwd, sc.cam["image_width"], sc.cam["image_height"])
        del fwd
        emit(f"== {name} ({prim}): P={sc.P} {sc.cam['image_width']}x{sc.cam['image_height']} gamma={sc.gamma} rich={sc.rich_info}  [{time.time() - t0:.0f}s]")
        emit("   work:", json.dumps(ws))
        res = {"work": ws, "ints": {}, "err": {}}
        for who in ("exact", "fastA"):
            ints = {k: mismatch_count(runs[who][k], base[k]) for k in harness.INT_KEYS if k in base and k in runs[who]}
            res["ints"][who] = ints
            emit(f"   {who:6s} integer mismatches vs refA: {ints}")
        pairs = [("refB", "refA"), ("exact", "refA"), ("fastA", "refA"), ("fastB", "fastA")]
        if "truth" in runs:
            pairs += [("refA", "truth"), ("exact", "truth"), ("fastA", "truth")]
        emit("   %-13s " % "tensor" + " ".join("%-46s" % f"{a}~{b}: max p99.99 p99 med" for a, b in pairs))
        for k in FLOAT_KEYS:
            row = []
            for a, b in pairs:
                if a in runs and b in runs and k in runs[a] and k in runs[b]:
                    st = err_stats(runs[a][k], runs[b][k])
                    res["err"].setdefault(k, {})[f"{a}~{b}"] = st
                    row.append("%-46s" % ("%.1e %.1e %.1e %.1e" % (st["max"], st["p9999"], st["p99"], st["med"])) if st else "-")
                else:
                    row.append("%-46s" % "-")
            emit("   %-13s " % k + " ".join(row))
        allres[name] = res
        del runs
        torch.cuda.empty_cache()
    p = os.path.join(ROOT, "gpurun_out", "scale_report.json")
    prev = json.load(open(p)) if os.path.exists(p) else {}
    prev.update(allres)
    json.dump(prev, open(p, "w"), indent=1)
    txt.close()


if __name__ == "__main__":
    main()
