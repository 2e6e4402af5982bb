#!/usr/bin/env python3
"""Build the UNMODIFIED reference 2D rasterizer into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE -- nothing under triangle_splatting_b200/ may import or link this.

The reference (GaodeRender/triangle-splatting, submodules/diff-triangle-rasterization-2D) has no
CPU implementation of the path; its CUDA extension is the only executable ground truth.  This
recipe compiles the reference's own five source files *where they lie* under /root/reference
(src/rasterizer.cu, src/forward.cu, src/backward.cu, src/extension_interface.cu, ext.cpp --
the list in the reference's setup.py:13-19) with plain nvcc, using the same flags
torch.utils.cpp_extension would pass, for sm_100 (what TORCH_CUDA_ARCH_LIST=10.0 emits).
It does not run the reference's build system and copies no reference source into this repo:
the only output is oracle/_ref/ts2d_ref_C*.so (git-ignored, but shipped to the GPU box by gpurun).

The pybind module is named ``ts2d_ref_C`` (-DTORCH_EXTENSION_NAME) and exposes the reference's two
entry points ``rasterize_triangles`` / ``rasterize_triangles_backward`` (ext.cpp:4-9).

/root/reference does not exist on the GPU box, so this script is a no-op there (prebuilt .so is used).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF_ROOT = Path(os.environ.get("TS2D_REFERENCE_ROOT", "/root/reference")) / "submodules"
SOURCES = ["src/rasterizer.cu", "src/forward.cu", "src/backward.cu", "src/extension_interface.cu", "ext.cpp"]
# which -> (reference directory, pybind module name).  "3D" is the reference's second per-pixel primitive
# (submodules/diff-triangle-rasterization-3D: same host code and API, ray/plane barycentrics in view space).
VARIANTS = {"2D": ("diff-triangle-rasterization-2D", "ts2d_ref_C"), "3D": ("diff-triangle-rasterization-3D", "ts3d_ref_C")}
REF = REF_ROOT / VARIANTS["2D"][0]
MODNAME = VARIANTS["2D"][1]


def so_path(which: str = "2D") -> Path:
    return OUT / f"{VARIANTS[which][1]}{sysconfig.get_config_var('EXT_SUFFIX')}"


def available(which: str = "2D") -> bool:
    return so_path(which).exists()


def build(force: bool = False, verbose: bool = True, which: str = "2D") -> Path | None:
    REF = REF_ROOT / VARIANTS[which][0]  # noqa: N806
    MODNAME = VARIANTS[which][1]  # noqa: N806
    target = so_path(which)
    if not REF.exists():
        if verbose:
            print(f"[oracle/build_ref] {REF} not present; using prebuilt {target if target.exists() else '(none)'}")
        return target if target.exists() else None
    srcs = [REF / s for s in SOURCES]
    hdrs = list((REF / "src").glob("*.h"))
    if target.exists() and not force:
        newest = max(p.stat().st_mtime for p in srcs + hdrs + [Path(__file__)])
        if target.stat().st_mtime >= newest:
            return target
    import torch  # noqa: F401  (only needed for include / lib paths)
    from torch.utils import cpp_extension as ce

    OUT.mkdir(parents=True, exist_ok=True)
    objdir = OUT / ("obj" if which == "2D" else "obj" + which)
    objdir.mkdir(exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = [
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-DTORCH_EXTENSION_NAME={MODNAME}",
        "-D_GLIBCXX_USE_CXX11_ABI=1",
        "-std=c++17",
    ]
    nvcc_flags = [
        "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
        "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
        "--expt-relaxed-constexpr", "--compiler-options", "-fPIC",
        # what TORCH_CUDA_ARCH_LIST=10.0 produces for the reference (it has no arch flags of its own,
        # setup.py:20): plain compute_100/sm_100, default -O3 device code, default -fmad=true.
        "-gencode=arch=compute_100,code=sm_100",
    ]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")

    def compile_one(src: Path) -> Path:
        obj = objdir / (src.name + ".o")
        if src.suffix == ".cu":
            cmd = [nvcc, "-c", str(src), "-o", str(obj)] + inc + common + nvcc_flags
        else:
            cmd = ["g++", "-c", str(src), "-o", str(obj), "-O3", "-fPIC"] + inc + common
        if verbose:
            print("[oracle/build_ref]", " ".join(cmd[:4]), "...", flush=True)
        subprocess.run(cmd, check=True, cwd=str(REF))
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(compile_one, srcs))
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", str(target)] + [str(o) for o in objs]
    for d in libdirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    subprocess.run(link, check=True)
    if verbose:
        print(f"[oracle/build_ref] built {target}")
    return target


def load(which: str = "2D"):
    """Import the reference pybind module (needs torch imported first). Returns module or None."""
    MODNAME = VARIANTS[which][1]  # noqa: N806
    p = so_path(which)
    if not p.exists():
        return None
    import importlib.util
    import torch  # noqa: F401

    spec = importlib.util.spec_from_file_location(MODNAME, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    for w in ("2D", "3D"):
        print(build(force="--force" in sys.argv, which=w))
