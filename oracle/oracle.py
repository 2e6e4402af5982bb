"""ctypes/numpy front-end of the CPU oracle (oracle/ts2d_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ts2d_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py (cpu_baseline / --impl reference legs) may import this module.

``Oracle("f32")`` is the fp32 mirror of the reference arithmetic, ``Oracle("f64")`` the fp64 truth.
``forward`` / ``backward`` take and return numpy arrays with the shapes of the reference's pybind
entry points (R2D/src/extension_interface.h:7-62) plus every intermediate the reference keeps in its
opaque state buffers (R2D/src/param_struct.h:44-123), so parity tests can compare stage by stage.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
SRC = HERE / "ts2d_oracle.c"
TILE = 16


def lib_path(kind: str) -> Path:
    return BUILD / f"libts2d_oracle_{kind}.so"


def build(force: bool = False) -> None:
    """gcc the C restatement twice (REAL=float / REAL=double). No FMA contraction, no fast-math."""
    BUILD.mkdir(exist_ok=True)
    for kind, real, extra in (("f32", "float", []), ("f64", "double", []), ("f64d", "double", ["-DDECIDE_F32"])):
        out = lib_path(kind)
        if out.exists() and not force and out.stat().st_mtime >= SRC.stat().st_mtime:
            continue
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math", f"-DREAL={real}", *extra,
               str(SRC), "-o", str(out), "-lm"]
        subprocess.run(cmd, check=True)


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self, kind: str = "f32"):
        assert kind in ("f32", "f64", "f64d")  # f64d: fp64 values, the reference's fp32 decisions (header of ts2d_oracle.c)
        build()
        self.kind = kind
        self.real = np.float32 if kind == "f32" else np.float64
        self.creal = ctypes.c_float if kind == "f32" else ctypes.c_double
        self.lib = ctypes.CDLL(str(lib_path(kind)))
        assert self.lib.ts2d_oracle_sizeof_real() == np.dtype(self.real).itemsize
        self.lib.ts2d_oracle_bin.restype = ctypes.c_int64

    def _r(self, a, shape=None) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(a, dtype=self.real))
        if shape is not None:
            a = a.reshape(shape)
        return a

    # ------------------------------------------------------------------ forward
    def forward(self, *, image_width, image_height, tanfovx, tanfovy, viewmatrix, projmatrix, campos, sh_degree, gamma,
                background_depth, background, vertex, shs, feature, opacity, back_culling=False, rich_info=False,
                scale_modifier=1.0, debug=False, stages="all", tile_step=1, tile_offset=0, primitive="2D", forced=None) -> dict:
        """primitive="3D": the reference's diff_triangle_rasterization_3D package (ts3d_oracle_* in ts2d_oracle.c).
        forced (kind "f64d" only): dict with the fp32 per-triangle state, lists and n_contrib of a GPU run of the reference (or of
        our bit-identical state): radii, clamped, v2d/area2/v_depth or v_view, normal_view, rgb, point_list, ranges, n_contrib.
        The composite then evaluates exactly that computation in double, taking every decision the way fp32 takes it."""
        assert primitive in ("2D", "3D")
        assert (forced is not None) == (self.kind == "f64d"), "the f64d oracle needs (and only it takes) a forced fp32 state"
        W, H = int(image_width), int(image_height)
        vertex = self._r(vertex)
        P = vertex.shape[0]
        shs_a = self._r(shs) if shs is not None and np.size(shs) else np.zeros((0,), self.real)
        feat_a = self._r(feature) if feature is not None and np.size(feature) else np.zeros((0,), self.real)
        # extension_interface.cu:44-50
        use_shs = feat_a.ndim <= 1 or (feat_a.shape[0] == 0 and shs_a.shape[0] > 0)
        C = 3 if use_shs else feat_a.shape[1]
        M = shs_a.shape[1] if shs_a.ndim == 3 and shs_a.shape[0] != 0 else 0
        vm, pm, cp = self._r(viewmatrix).reshape(16), self._r(projmatrix).reshape(16), self._r(campos).reshape(3)
        opacity = self._r(opacity).reshape(P)
        bg = self._r(background).reshape(C)
        real = self.real
        st = dict(W=W, H=H, P=P, C=C, M=M, D=int(sh_degree), use_shs=bool(use_shs), rich_info=bool(rich_info), gamma=float(gamma),
                  tanfovx=float(tanfovx), tanfovy=float(tanfovy), viewmatrix=vm, projmatrix=pm, campos=cp, vertex=vertex, shs=shs_a,
                  opacity=opacity, background=bg, background_depth=float(background_depth), back_culling=bool(back_culling),
                  primitive=primitive)
        st["radii"] = np.zeros(P, np.int32)
        st["v2d"] = np.zeros((P, 3, 2), real)
        st["area2"] = np.zeros(P, real)
        st["normal_view"] = np.zeros((P, 3), real)
        st["v_depth"] = np.zeros((P, 3), real)
        st["depth"] = np.zeros(P, real)
        st["rgb"] = np.zeros((P, 3), real)
        st["clamped"] = np.zeros((P, 3), np.uint8)
        st["tiles_touched"] = np.zeros(P, np.uint32)
        st["rect_min"] = np.zeros((P, 2), np.uint32)
        st["rect_max"] = np.zeros((P, 2), np.uint32)
        cr = self.creal
        if primitive == "3D":
            st["v_view"] = np.zeros((P, 3, 3), real)
            if P > 0:
                self.lib.ts3d_oracle_preprocess(
                    W, H, P, int(sh_degree), M, int(bool(use_shs)), int(bool(back_culling)), _p(vm), _p(pm), _p(cp), _p(vertex), _p(shs_a),
                    _p(st["radii"]), _p(st["v_view"]), _p(st["normal_view"]), _p(st["depth"]), _p(st["rgb"]), _p(st["clamped"]),
                    _p(st["tiles_touched"]), _p(st["rect_min"]), _p(st["rect_max"]))
        elif P > 0:
            self.lib.ts2d_oracle_preprocess(
                W, H, P, int(sh_degree), M, int(bool(rich_info)), int(bool(use_shs)), int(bool(back_culling)), cr(tanfovx), cr(tanfovy),
                _p(vm), _p(pm), _p(cp), _p(vertex), _p(shs_a), _p(st["radii"]), _p(st["v2d"]), _p(st["area2"]), _p(st["normal_view"]),
                _p(st["v_depth"]), _p(st["depth"]), _p(st["rgb"]), _p(st["clamped"]), _p(st["tiles_touched"]), _p(st["rect_min"]),
                _p(st["rect_max"]))
        if forced is not None and P > 0:
            for k in ("v_view",) if primitive == "3D" else ("v2d", "area2", "v_depth"):
                st[k] = np.ascontiguousarray(np.asarray(forced[k], dtype=real).reshape(st[k].shape))
            st["normal_view"] = np.ascontiguousarray(np.asarray(forced["normal_view"], dtype=real).reshape(P, 3))
            if use_shs:
                st["rgb"] = np.ascontiguousarray(np.asarray(forced["rgb"], dtype=real).reshape(P, 3))
            st["radii"] = np.ascontiguousarray(np.asarray(forced["radii"], dtype=np.int32))
            st["clamped"] = np.ascontiguousarray(np.asarray(forced["clamped"], dtype=np.uint8).reshape(P, 3))
        st["feature"] = st["rgb"] if use_shs else feat_a.reshape(P, C)
        if stages == "preprocess":
            return st
        gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
        # the reference sorts the fp32 bit pattern of depth (rasterizer.cu:69)
        depth32 = np.ascontiguousarray(st["depth"].astype(np.float32)).view(np.uint32)
        st["point_offsets"] = np.zeros(P, np.uint32)
        R = int(self.lib.ts2d_oracle_bin(W, H, P, _p(st["tiles_touched"]), _p(st["rect_min"]), _p(st["rect_max"]), _p(depth32),
                                         _p(st["point_offsets"]), None, None, None)) if P > 0 else 0
        st["num_rendered"] = R
        st["keys"] = np.zeros(R, np.uint64)
        st["point_list"] = np.zeros(R, np.uint32)
        st["ranges"] = np.zeros((gx * gy, 2), np.uint32)
        if R > 0:
            self.lib.ts2d_oracle_bin(W, H, P, _p(st["tiles_touched"]), _p(st["rect_min"]), _p(st["rect_max"]), _p(depth32),
                                     _p(st["point_offsets"]), _p(st["keys"]), _p(st["point_list"]), _p(st["ranges"]))
        if forced is not None:
            st["num_rendered"] = int(np.asarray(forced["point_list"]).size)
            st["point_list"] = np.ascontiguousarray(np.asarray(forced["point_list"], dtype=np.uint32))
            st["ranges"] = np.ascontiguousarray(np.asarray(forced["ranges"], dtype=np.uint32).reshape(gx * gy, 2))
        if stages == "bin":
            return st
        st["final_T"] = np.ones((H, W), real)
        st["n_contrib"] = (np.ascontiguousarray(np.asarray(forced["n_contrib"], dtype=np.uint32).reshape(H, W)) if forced is not None
                           else np.zeros((H, W), np.uint32))
        st["out_feature"] = np.zeros((C, H, W), real)
        st["out_depth"] = np.zeros((H, W), real)
        st["out_normal"] = np.zeros((3, H, W), real)
        st["contrib_sum"] = np.zeros(P, real)
        st["contrib_max"] = np.zeros(P, real)
        if P == 0:  # extension_interface.cu:130: zero outputs
            return st
        st["tile_step"], st["tile_offset"] = int(tile_step), int(tile_offset)
        if primitive == "3D":
            self.lib.ts3d_oracle_render(
                W, H, C, cr(gamma), int(bool(rich_info)), cr(tanfovx), cr(tanfovy), _p(st["ranges"]), _p(st["point_list"]), _p(st["v_view"]),
                _p(st["normal_view"]), _p(np.ascontiguousarray(st["feature"])), _p(opacity), cr(background_depth), _p(bg), _p(st["final_T"]),
                _p(st["n_contrib"]), _p(st["out_feature"]), _p(st["out_depth"]), _p(st["out_normal"]), _p(st["contrib_sum"]),
                _p(st["contrib_max"]), P, int(tile_step), int(tile_offset))
            return st
        self.lib.ts2d_oracle_render(
            W, H, C, cr(gamma), int(bool(rich_info)), _p(st["ranges"]), _p(st["point_list"]), _p(st["v2d"]), _p(st["area2"]),
            _p(st["normal_view"]), _p(st["v_depth"]), _p(np.ascontiguousarray(st["feature"])), _p(opacity), cr(background_depth), _p(bg),
            _p(st["final_T"]), _p(st["n_contrib"]), _p(st["out_feature"]), _p(st["out_depth"]), _p(st["out_normal"]), _p(st["contrib_sum"]),
            _p(st["contrib_max"]), P, int(tile_step), int(tile_offset))
        st["tile_step"], st["tile_offset"] = int(tile_step), int(tile_offset)
        return st

    # ------------------------------------------------------------------ backward
    def backward(self, st: dict, dL_dout_feature, dL_dout_depth=None, dL_dout_normal=None) -> dict:
        W, H, P, C, M = st["W"], st["H"], st["P"], st["C"], st["M"]
        real, cr = self.real, self.creal
        rich = st["rich_info"]
        g_img = self._r(dL_dout_feature).reshape(C, H, W)
        g_dep = self._r(dL_dout_depth).reshape(H, W) if rich and dL_dout_depth is not None else np.zeros((H, W), real)
        g_nrm = self._r(dL_dout_normal).reshape(3, H, W) if rich and dL_dout_normal is not None else np.zeros((3, H, W), real)
        out = dict(g_v2d=np.zeros((P, 3, 2)), g_normal=np.zeros((P, 3)), g_vdepth=np.zeros((P, 3)), g_feature=np.zeros((P, C)),
                   g_opacity=np.zeros(P))
        out["dL_dvertex"] = np.zeros((P, 3, 3), real)
        out["dL_dcenter2D"] = np.zeros((P, 2), real)
        out["dL_dshs"] = np.zeros((P, M, 3), real)
        if P == 0:
            out["dL_dfeature"] = np.zeros((P, C), real)
            out["dL_dopacity"] = np.zeros((P, 1), real)
            return out
        feat = np.ascontiguousarray(st["feature"])
        if st.get("primitive", "2D") == "3D":
            out["g_vview"] = np.zeros((P, 3, 3))
            self.lib.ts3d_oracle_render_bwd(
                W, H, C, cr(st["gamma"]), int(rich), cr(st["tanfovx"]), cr(st["tanfovy"]), _p(st["ranges"]), _p(st["point_list"]),
                _p(st["v_view"]), _p(st["normal_view"]), _p(feat), _p(st["opacity"]), cr(st["background_depth"]), _p(st["background"]),
                _p(st["final_T"]), _p(st["n_contrib"]), _p(g_img), _p(g_dep), _p(g_nrm), P, _p(out["g_vview"]), _p(out["g_normal"]),
                _p(out["g_feature"]), _p(out["g_opacity"]), int(st.get("tile_step", 1)), int(st.get("tile_offset", 0)))
            g_rgb = out["g_feature"] if st["use_shs"] else np.zeros((P, 3))
            self.lib.ts3d_oracle_preprocess_bwd(
                P, st["D"], M, int(st["use_shs"]), _p(st["viewmatrix"]), _p(st["campos"]), _p(st["vertex"]), _p(st["shs"]), _p(st["radii"]),
                _p(st["clamped"]), _p(st["v_view"]), _p(out["g_vview"]), _p(out["g_normal"]), _p(np.ascontiguousarray(g_rgb)),
                _p(out["dL_dvertex"]), _p(out["dL_dcenter2D"]), _p(out["dL_dshs"]))
            out["dL_dfeature"] = out["g_feature"].astype(real)
            out["dL_dopacity"] = out["g_opacity"].astype(real).reshape(P, 1)
            return out
        self.lib.ts2d_oracle_render_bwd(
            W, H, C, cr(st["gamma"]), int(rich), _p(st["ranges"]), _p(st["point_list"]), _p(st["v2d"]), _p(st["area2"]), _p(st["normal_view"]),
            _p(st["v_depth"]), _p(feat), _p(st["opacity"]), cr(st["background_depth"]), _p(st["background"]), _p(st["final_T"]),
            _p(st["n_contrib"]), _p(g_img), _p(g_dep), _p(g_nrm), P, _p(out["g_v2d"]), _p(out["g_normal"]), _p(out["g_vdepth"]),
            _p(out["g_feature"]), _p(out["g_opacity"]), int(st.get("tile_step", 1)), int(st.get("tile_offset", 0)))
        g_rgb = out["g_feature"] if st["use_shs"] else np.zeros((P, 3))
        self.lib.ts2d_oracle_preprocess_bwd(
            W, H, P, st["D"], M, int(st["use_shs"]), int(rich), cr(st["tanfovx"]), cr(st["tanfovy"]), _p(st["viewmatrix"]),
            _p(st["projmatrix"]), _p(st["campos"]), _p(st["vertex"]), _p(st["shs"]), _p(st["radii"]), _p(st["clamped"]),
            _p(out["g_v2d"]), _p(out["g_normal"]), _p(out["g_vdepth"]), _p(np.ascontiguousarray(g_rgb)), _p(out["dL_dvertex"]),
            _p(out["dL_dcenter2D"]), _p(out["dL_dshs"]))
        # in SH mode the reference returns its dL_drgb scratch as dL_dfeature (extension_interface.cu:232)
        out["dL_dfeature"] = out["g_feature"].astype(real)
        out["dL_dopacity"] = out["g_opacity"].astype(real).reshape(P, 1)
        return out


if __name__ == "__main__":
    build(force=True)
    print("built", lib_path("f32"), lib_path("f64"))
