"""In-tree build of libgcb200.so (the C-ABI library) for sm_100a.

`python -m mpc_b200.build` or `mpc_b200.build.build()`; nvcc cross-compiles
without a GPU.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libgcb200.so")
SOURCES = ["gcb200.cu", "gc_garble.cu", "gc_eval.cu", "plan.cpp", "circuit_host.cpp"]
HEADERS = ["aes_core.cuh", "gc_kernels.cuh", "ot_kernels.cuh", "stream_kernels.cuh", "plan.hpp", "async.hpp",
           "gc_launch.hpp", "../../include/gcb200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--use_fast_math",
    "-Xptxas", "-v", "-diag-suppress", "128",
]
OBJ_DIR = os.path.join(HERE, "build")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit to an object (in parallel: the gate kernels are two
    units of 36 instantiations each), then link the shared library."""
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("GCB_NVCC_EXTRA", "").split()
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for f in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.splitext(f)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, f)]
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    out, failed = [], False
    for cmd, obj, proc in jobs:
        text, _ = proc.communicate()
        out.append(" ".join(cmd) + "\n" + text)
        failed |= proc.returncode != 0
    if not failed:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + [obj for _, obj, _ in jobs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        out.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        failed = r.returncode != 0
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write("\n".join(out))
    if verbose or failed:
        sys.stderr.write("\n".join(out))
    if failed:
        raise RuntimeError(f"nvcc failed; see {log}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
