"""Build libts2d.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m triangle_splatting_b200.build [--force]

Output: triangle_splatting_b200/lib/libts2d.so (git-ignored; travels to the GPU box with gpurun).
No torch / pybind dependency: the library is plain CUDA runtime (no CUB / Thrust either: the sorts and scans are hand-written, ts2d_sort.cuh).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libts2d.so"
SOURCES = ["ts2d_api.cu", "ts2d_preprocess.cu", "ts2d_binning.cu", "ts2d_render_fwd.cu", "ts2d_render_bwd.cu", "ts2d_render_fwd_fast.cu", "ts2d_render_bwd_fast.cu", "ts2d_bwd_reduce.cu", "ts2d_exchange.cu", "ts2d_prim3d.cu", "ts2d_prim3d_fast.cu", "ts2d_loss.cu", "ts2d_geometry_loss.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# NOTE: no --use_fast_math: bit-exact tile binning needs IEEE div/sqrt and the default FMA contraction.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + ARCH


def nvcc() -> str:
    return os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ts2d.h", Path(__file__)]
    return LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)

    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ts2d.h", Path(__file__)]
    newest_header = max(p.stat().st_mtime for p in headers)
    extra = os.environ.get("TS2D_NVCC_EXTRA", "")
    stamp = objdir / "flags.txt"
    flags_changed = (not stamp.exists()) or stamp.read_text() != extra
    stamp.write_text(extra)

    def one(src: str) -> Path:
        obj = objdir / (src + ".o")
        if not force and not flags_changed and obj.exists() and obj.stat().st_mtime > max(newest_header, (CSRC / src).stat().st_mtime):
            return obj  # up to date: only the translation units that changed are recompiled
        # TS2D_NVCC_EXTRA: extra defines for tuning experiments (e.g. "-DTS2D_FWD_MINB=4"); empty for the shipped build
        cmd = [nvcc(), "-c", str(CSRC / src), "-o", str(obj)] + NVCC_FLAGS + os.environ.get("TS2D_NVCC_EXTRA", "").split()
        r = subprocess.run(cmd, capture_output=True, text=True)
        (objdir / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [nvcc(), "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ARCH + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


# ---------------------------------------------------------------------------------------------- pybind11 layer (_C_native)
NATIVE_SRC = CSRC / "ts2d_pybind.cpp"
NATIVE_NAME = "_C_native"


def native_path() -> Path:
    import sysconfig

    return PKG / f"{NATIVE_NAME}{sysconfig.get_config_var('EXT_SUFFIX')}"


def native_needs_build() -> bool:
    t = native_path()
    if not t.exists():
        return True
    deps = [NATIVE_SRC, PKG.parent / "include" / "ts2d.h"]
    return t.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """g++ build of csrc/ts2d_pybind.cpp against torch's headers into triangle_splatting_b200/_C_native*.so (git-ignored, in-tree so
    that it travels to the GPU box); links libts2d.so through an $ORIGIN-relative rpath.  No CUDA code in it: nvcc is not involved."""
    target = native_path()
    if not force and not native_needs_build():
        return target
    build(force=False)
    import sysconfig

    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{PKG.parent / 'include'}"]
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-DTORCH_EXTENSION_NAME={NATIVE_NAME}",
           "-D_GLIBCXX_USE_CXX11_ABI=1", str(NATIVE_SRC), "-o", str(target)] + inc
    for d in ce.library_paths("cuda"):
        cmd += [f"-L{d}", f"-Wl,-rpath,{d}"]
    cmd += [f"-L{LIBDIR}", "-Wl,-rpath,$ORIGIN/lib", "-lts2d", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {NATIVE_SRC.name}:\n{r.stdout}\n{r.stderr}")
    return target


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
