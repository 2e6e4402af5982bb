"""CPU tests of the oracle itself: golden vectors from the reference's CUDA build, finite-difference
gradient check (fp64), fp32-vs-fp64 agreement, structural properties, edge cases."""
import os

import numpy as np
import pytest
import torch

import harness
from harness import GRAD_KEYS, IMAGE_KEYS, assert_close_modulo_flips, mismatch_count
from oracle.oracle import Oracle
from triangle_splatting_b200.scenes import make_scene


def _golden(name, primitive="2D"):
    p = harness.golden_path(name, primitive)
    if not os.path.exists(p):
        pytest.skip(f"golden fixture {p} missing")
    return dict(np.load(p))


@pytest.mark.parametrize("name", list(harness.GOLDEN_SCENES))
def test_oracle_vs_reference_golden(name):
    """Pins the oracle against the reference's own outputs (B200, oracle/_ref).  The fp32 mirror has no FMA
    contraction and uses glibc powf/expf, so integer fields get a tiny mismatch budget (decision boundaries)
    and floats a looser tolerance than the GPU-vs-GPU bar."""
    sc = harness.golden_scene(name)
    gold = _golden(name)
    orc = harness.run_oracle(sc, "f32")
    P = sc.P
    budget = max(1, P // 1000)
    assert mismatch_count(orc["radii"], gold["radii"]) <= budget
    assert mismatch_count(orc["tiles_touched"], gold["tiles_touched"]) <= budget
    assert abs(int(orc["num_rendered"]) - int(gold["num_rendered"])) <= 4 * budget
    if int(orc["num_rendered"]) == int(gold["num_rendered"]):
        assert mismatch_count(orc["point_list"], gold["point_list"]) <= 8 * budget
        assert mismatch_count(orc["ranges"], gold["ranges"]) <= 8 * budget
    if mismatch_count(orc["point_list"], gold["point_list"]) == 0:
        bad = mismatch_count(orc["n_contrib"], gold["n_contrib"])
        assert bad <= max(2, orc["n_contrib"].size // 2000), f"n_contrib differs on {bad} pixels"
        for k in IMAGE_KEYS + GRAD_KEYS:
            if k in gold and k in orc:
                assert_close_modulo_flips(orc[k], gold[k], f"golden[{name}].{k}")
    o64 = harness.run_oracle(sc, "f64")
    for k in ("out_feature",) + GRAD_KEYS:
        if k in gold:
            assert_close_modulo_flips(gold[k], o64[k], f"golden-vs-f64[{name}].{k}")


@pytest.mark.parametrize("name", list(harness.GOLDEN_SCENES_3D))
def test_oracle3d_vs_reference_golden(name):
    """Same pinning for the 3D primitive: tests/golden/3d_*.npz are outputs of the reference's own
    diff_triangle_rasterization_3D build (oracle/_ref/ts3d_ref_C*.so) on a B200."""
    sc = harness.golden_scene(name, "3D")
    gold = _golden(name, "3D")
    orc = harness.run_oracle(sc, "f32", primitive="3D")
    budget = max(1, sc.P // 1000)
    assert mismatch_count(orc["radii"], gold["radii"]) <= budget
    assert mismatch_count(orc["tiles_touched"], gold["tiles_touched"]) <= budget
    assert abs(int(orc["num_rendered"]) - int(gold["num_rendered"])) <= 4 * budget
    if mismatch_count(orc["point_list"], gold["point_list"]) == 0:
        bad = mismatch_count(orc["n_contrib"], gold["n_contrib"])
        assert bad <= max(2, orc["n_contrib"].size // 2000), f"n_contrib differs on {bad} pixels"
        for k in IMAGE_KEYS + GRAD_KEYS:
            if k in gold and k in orc:
                assert_close_modulo_flips(orc[k], gold[k], f"golden3d[{name}].{k}")


@pytest.mark.parametrize("primitive", ["2D", "3D"])
def test_fd_gradients_fp64(primitive):
    """With opacity == 1 the 3D backward's G < 1/255 test coincides with the forward's alpha < 1/255 test, so the
    reference's 3D gradients ARE the derivative of its forward (see the quirk list in csrc/ts2d_prim3d.cu)."""
    sc = make_scene("t", 200, 48, 40, sh_degree=2, rich_info=True, geometry_grads=True, seed=5, rho_px=4.0)
    if primitive == "3D":
        sc.opacity.fill_(1.0)  # alpha clamps at 0.99 near the centre: dL_dpower = 0 there on both sides
    o = Oracle("f64")
    kw = sc.settings_kwargs()
    kw.pop("debug")
    kw = {k: (v.numpy().astype(np.float64) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    V, S, O = (t.numpy().astype(np.float64) for t in (sc.vertex, sc.shs, sc.opacity))
    gf, gd, gn = (sc.grads[k].numpy().astype(np.float64) for k in ("dL_dout_feature", "dL_dout_depth", "dL_dout_normal"))

    def loss(V, S, O):
        st = o.forward(**kw, vertex=V, shs=S, feature=None, opacity=O, primitive=primitive)
        return (st["out_feature"] * gf).sum() + (st["out_depth"] * gd).sum() + (st["out_normal"] * gn).sum(), st

    _, st = loss(V, S, O)
    g = o.backward(st, gf, gd, gn)
    rng = np.random.default_rng(0)
    vis = np.nonzero(st["radii"] > 0)[0]
    checks = (("V", V, g["dL_dvertex"]), ("S", S, g["dL_dshs"])) + ((("O", O, g["dL_dopacity"]),) if primitive == "2D" else ())
    for key, arr, grad in checks:
        for _ in range(6):
            i = rng.choice(vis)
            idx = (i,) + tuple(rng.integers(0, s) for s in arr.shape[1:])
            h = 1e-6 * max(1.0, abs(arr[idx]))
            args = dict(V=V, S=S, O=O)
            a, b = arr.copy(), arr.copy()
            a[idx] += h
            b[idx] -= h
            args[key] = a
            lp, _ = loss(**args)
            args[key] = b
            lm, _ = loss(**args)
            fd = (lp - lm) / (2 * h)
            assert fd == pytest.approx(grad[idx], rel=2e-4, abs=1e-9)


def test_fp32_mirror_agrees_with_fp64_truth():
    sc = make_scene("t", 1500, 96, 80, sh_degree=3, rich_info=True, geometry_grads=True, seed=7)
    a, b = harness.run_oracle(sc, "f32"), harness.run_oracle(sc, "f64")
    assert mismatch_count(a["radii"], b["radii"]) <= 2
    if mismatch_count(a["point_list"], b["point_list"]) == 0:
        for k in ("out_feature", "depth", "normal") + GRAD_KEYS:
            assert_close_modulo_flips(a[k], b[k], k)


@pytest.mark.parametrize("primitive", ["2D", "3D"])
def test_structure_and_edge_cases(primitive):
    o = Oracle("f32")
    sc = make_scene("t", 800, 75, 37, sh_degree=0, seed=3, rho_px=6.0)  # ragged image
    r = harness.run_oracle(sc, "f32", primitive=primitive)
    keys = r["keys"]
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    pl = r["point_list"].astype(np.int64)
    assert np.all(pl[1:][same] > pl[:-1][same])
    assert int(r["num_rendered"]) == int(r["tiles_touched"].sum())
    rng = r["ranges"].astype(np.int64)
    assert int((rng[:, 1] - rng[:, 0]).sum()) == int(r["num_rendered"])
    assert np.all(r["final_T"] <= 1.0) and np.all(r["final_T"] >= 0.0)
    # empty input
    sc0 = make_scene("e", 0, 32, 32)
    r0 = harness.run_oracle(sc0, "f32", primitive=primitive)
    assert int(r0["num_rendered"]) == 0 and r0["out_feature"].shape == (3, 32, 32)
    # all culled: image == background
    sc1 = make_scene("b", 100, 32, 32, seed=2)
    sc1.vertex[..., 2] -= 100.0
    r1 = harness.run_oracle(sc1, "f32", primitive=primitive)
    assert int(r1["num_rendered"]) == 0
    assert np.array_equal(r1["out_feature"], np.broadcast_to(sc1.background.numpy()[:, None, None], (3, 32, 32)))


@pytest.mark.parametrize("primitive,name", [("2D", n) for n in harness.GOLDEN_SCENES] + [("3D", n) for n in harness.GOLDEN_SCENES_3D])
def test_truth_given_reference_decisions_vs_golden(primitive, name):
    """Oracle kind "f64d" (fp64 values, the reference's fp32 decisions and stopping points) on the golden fixtures: with the
    decisions pinned, the reference's fp32 pixels are plain roundings of the truth -- no flips, so a MAX-norm bar holds -- and
    the reference's gradients sit where the parity reports put them (1e-2 class in this metric, not 1e-5)."""
    sc = harness.golden_scene(name, primitive)
    gold = _golden(name, primitive)
    truth = harness.run_truth(sc, gold, primitive, backward="dL_dvertex" in gold)
    if primitive == "2D":  # the shadow reproduces the 2D decisions exactly (same contraction as the reference's SASS): max-norm bar
        assert harness.rel_err(gold["out_feature"], truth["out_feature"]) <= 2e-5
        assert harness.rel_err(gold["final_T"], truth["final_T"], rel_floor=1e-2) <= 1e-3
        if sc.rich_info:
            assert harness.rel_err(gold["depth"], truth["depth"]) <= 2e-5
    else:  # the 3D shadow is plain fp32 (the reference's contraction of its ray / plane arithmetic is not restated): a pair within
        # an ulp of a threshold may flip -> quantile bar, coarse max
        q = harness.err_quantiles(gold["out_feature"], truth["out_feature"])
        assert q[1] <= 2e-5 and q[2] <= 1e-3 and q[3] <= 1e-2, f"out_feature quantiles {q}"
    if sc.rich_info and primitive == "2D":
        assert harness.rel_err(gold["contrib_sum"], truth["contrib_sum"]) <= 5e-4  # an fp32 atomic sum of up to thousands of terms
        assert harness.rel_err(gold["contrib_max"], truth["contrib_max"]) <= 1e-3  # alpha * T, T a product of hundreds of (1 - alpha) factors
    # forced stopping points are consistent with the transmittance the double arithmetic sees
    rng = gold["ranges"].astype(np.int64)
    gx = (sc.cam["image_width"] + 15) // 16
    tile_of_pix = (np.arange(sc.cam["image_height"])[:, None] // 16) * gx + (np.arange(sc.cam["image_width"])[None, :] // 16)
    stopped = gold["n_contrib"] < (rng[:, 1] - rng[:, 0])[tile_of_pix]
    assert np.all(truth["final_T"][stopped] <= 1.0001e-4)
    if "dL_dvertex" in gold:
        for k in GRAD_KEYS:
            if k in gold:
                q = harness.err_quantiles(gold[k], truth[k])
                assert q[1] <= 2e-3 and q[3] <= 0.5, f"{k}: reference vs truth quantiles {q}"
