"""
Builds libpmm_b200.so IN-TREE with nvcc for sm_100a (no JIT cache: the .so travels with the repo
snapshot to the GPU box).  `python -m polars_matmul_b200.build [--force]`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libpmm_b200.so")
SOURCES = ["pmm_prep.cu", "pmm_generic.cu", "pmm_merge.cu", "pmm_rescore.cu", "pmm_tc_kernels.cu", "pmm_stage.cu", "pmm_nccl.cu", "pmm_api.cu"]
HEADERS = ["pmm_common.cuh", "pmm_kernels.h", "pmm_tc.cuh", "pmm_stage.h", "pmm_nccl.h", os.path.join("..", "..", "include", "pmm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-fmad=false",            # no implicit FMA contraction: the metric pass mirrors the reference op for op
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-ccbin", "/usr/bin/g++",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpmm_b200 cannot be built (there is no CPU fallback)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs, todo = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            todo.append(cmd)
    if todo:   # the translation units are independent: compile them side by side (the tensor-core file dominates)
        from concurrent.futures import ThreadPoolExecutor
        jobs = jobs or min(len(todo), os.cpu_count() or 1)

        def run(cmd):
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)

        with ThreadPoolExecutor(max_workers=max(1, jobs)) as ex:
            list(ex.map(run, todo))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, "-ccbin", "/usr/bin/g++",
               "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
