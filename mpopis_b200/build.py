"""Builds mpopis_b200/libmpopis_b200.so (sm_100a only) with nvcc, in-tree.

The shared library is the product: a C-ABI (include/mpopis_b200.h) over hand-written CUDA kernels.
No torch extension machinery is involved; the .so has no Python, torch or NCCL link dependency
(NCCL is dlopen'ed at run time for the sharded configuration).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmpopis_b200.so"
OBJ = PKG / "csrc" / "_obj"
SOURCES = ["mpopis_b200.cu", "rollout.cu", "rollout_split.cu", "rollout_aux.cu", "sampling.cu", "stats.cu", "linalg.cu", "small_adapt.cu", "sort.cu", "select.cu", "comm.cu", "cma.cu"]
HEADERS = [CSRC / "engine.cuh", CSRC / "comm.cuh", CSRC / "chol_tile.cuh", CSRC / "car_model.cuh", CSRC / "rollout_kernels.cuh", PKG.parent / "include" / "mpopis_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    OBJ.mkdir(exist_ok=True)
    jobs = []
    for src in SOURCES:
        s, o = CSRC / src, OBJ / (src[:-3] + ".o")
        if force or _stale(o, [s, *HEADERS]):
            jobs.append([nvcc, *flags, "-c", str(s), "-o", str(o)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr, file=sys.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [str(OBJ / (s[:-3] + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
