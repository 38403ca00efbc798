"""In-tree build of libmrt.so (hand-written sm_100a CUDA kernels + the C ABI of include/mrt.h).

``python -m mri_raytracer_b200.build`` (or ``__graft_entry__.build()``) cross-compiles with
nvcc; no GPU is needed to build.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libmrt.so"
OBJ = PKG / "build"

SOURCES = ["c_api.cu", "forward.cu", "forward_tma.cu", "backward.cu", "occupancy.cu", "misc.cu", "slab.cu", "host_pipeline.cu", "inr.cu", "adaptive.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# No -use_fast_math: the parity contract needs IEEE div/sqrt and full-precision expf.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libmrt.so")


def _digest(extra: list[str]) -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [INCLUDE / "mrt.h"]):
        h.update(p.name.encode()); h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + ARCH + extra).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> Path:
    extra = list(extra_flags or [])
    if os.environ.get("MRT_NVCC_EXTRA"):
        extra += os.environ["MRT_NVCC_EXTRA"].split()
    stamp = OBJ / "stamp"
    dig = _digest(extra)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = OBJ / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *ARCH, *extra, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            (OBJ / (src + ".ptxas.log")).write_text(r.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", *ARCH, "-o", str(LIB), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
