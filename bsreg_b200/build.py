"""Build libbsreg_cuda.so in-tree with nvcc for sm_100a (no JIT, no torch extension).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "lib" / "libbsreg_cuda.so"
SOURCES = ["api.cu", "sweep.cu", "semsweep.cu", "logdet.cu", "slxdelta.cu", "chain.cu", "persist.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "bsreg_cuda.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, variant: str = "") -> Path:
    """variant: suffix of a side build (lib/libbsreg_cuda_<variant>.so, own object directory) for A/B profiling."""
    if not variant and not force and not needs_build():
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    objdir = ROOT / "lib" / ("obj_" + variant if variant else "obj")
    out = LIB.with_name(f"libbsreg_cuda_{variant}.so") if variant else LIB
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    extra = ["-DBSREG_TRACE"] if os.environ.get("BSREG_TRACE") else []   # clock64 trace points (diagnostics)

    def compile_one(src: str):
        obj = objdir / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        (objdir / (src[:-3] + ".ptxas.log")).write_text(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(out), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var[0] if var else ""))
