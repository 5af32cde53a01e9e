"""Builds libnrb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libnrb200.so")
SOURCES = ["engine.cu", "plan.cpp", "rrng.cpp", "host_simd.cpp"]
HEADERS = ["kernels.cuh", "device_views.cuh", "stream_kernels.cuh", "rank_kernels.cuh", "index_kernels.cuh", "philox.cuh", "plan.h", "handle.h", "index.cuh", "lists.cuh", "run.cuh", "group.cuh", "api.cuh",
           os.path.join("..", "..", "include", "nrb200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-fopenmp",
    "-shared", "-cudart", "static", "-lgomp",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libnrb200 cannot be built")
    return exe


def nvcc_available() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    rule = os.environ.get("NRB200_ZERO_WIDTH_RULE")  # development: flip THE ZERO-WIDTH RULE (csrc/device_views.cuh)
    cmd = [nvcc(), *NVCC_FLAGS, *([f"-DNRB_ZERO_WIDTH_RULE={int(rule)}"] if rule else []), *(["-Xptxas", "-v"] if verbose else []), "-o", SO,
           *[os.path.join(CSRC, f) for f in SOURCES]]
    env = dict(os.environ)
    # nvcc's host compiler: the system g++ (the image also exports CC/CXX wrappers)
    cmd[1:1] = ["-ccbin", shutil.which("g++", path="/usr/bin:/bin") or "g++"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libnrb200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
