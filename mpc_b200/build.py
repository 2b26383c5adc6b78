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
SOURCES = ["gcb200.cu", "stream.cu", "plan.cpp"]
HEADERS = ["aes_core.cuh", "gc_kernels.cuh", "ot_kernels.cuh", "stream_kernels.cuh", "plan.hpp", "hostpipe.hpp",
           "common.hpp", "../../include/gcb200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "--use_fast_math",
    "-Xptxas", "-v", "-diag-suppress", "128",
]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    extra = os.environ.get("GCB_NVCC_EXTRA", "").split()      # e.g. -DGCB_AES_TABLES=2 for experiments
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", SO] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError(f"nvcc failed ({r.returncode}); see {log}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
