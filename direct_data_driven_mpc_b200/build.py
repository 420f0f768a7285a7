"""Builds csrc/libddmpc.so for sm_100a with nvcc (in-tree, no JIT cache).

Each translation unit is compiled to an object file (in parallel, only when it or a header is newer than the object),
then linked into the shared library.  `python build.py --force -v` rebuilds everything and prints ptxas' resource usage.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["setup.cu", "solve.cu", "fast_loop.cu", "perloop_loop.cu", "cvx_loop.cu", "tc_loop.cu", "gemm_loop.cu", "dmma_loop.cu",
           "scenario_gen.cu", "probes.cu"]
HEADERS = ["common.cuh", "linalg.cuh", "plan.cuh", "fast_common.cuh", "ws_kernel.cuh",
           os.path.join("..", "..", "include", "ddmpc.h")]
LIB = os.path.join(CSRC, "libddmpc.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libddmpc.so cannot be built")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, src.replace(".cu", ".o"))


def _stale(src: str, hdr_t: float) -> bool:
    o = _obj(src)
    return (not os.path.exists(o)) or os.path.getmtime(o) < max(os.path.getmtime(os.path.join(CSRC, src)), hdr_t)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _newest_header()
    todo = [s for s in _sources() if force or _stale(s, hdr_t)]

    def compile_one(src: str) -> None:
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", _obj(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as pool:
        list(pool.map(compile_one, todo))
    link = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", LIB, *[_obj(s) for s in _sources()]]
    subprocess.run(link, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--force" in sys.argv))
