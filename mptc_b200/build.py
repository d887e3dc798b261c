"""Builds libmptc_b200.so (the CUDA kernels + the C-ABI of include/mptc_gpu.h) in-tree with
nvcc for sm_100a.  No torch extension machinery: the library has a plain C ABI."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("MPTC_LIB") or os.path.join(HERE, "libmptc_b200.so")   # MPTC_LIB: A/B builds (profiling)

SOURCES = ["mptc_kernels.cu", "mptc_inter.cu", "mptc_inter_wide.cu", "mptc_intra_rows.cu", "mptc_sparse.cu", "mptc_pixel.cu", "mptc_decode.cu", "mptc_capi.cu", "mptc_host.cpp"]
HEADERS = ["mptc_kernels.h", "mptc_device.cuh", "mptc_uniform_eval.cuh", "mptc_inter_tile.cuh", "mptc_host.h", os.path.join("..", "..", "include", "mptc_gpu.h"),
           os.path.join("..", "..", "include", "mptc_codec.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # belt and braces: every reference FP32 op is individually rounded
    "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-pthread",
    "-shared", "-lcudart",
]


# Sources that determine each search kernel's machine code: profiles/ncu_counters.json is stamped with
# their hash (profiles/ncu_summary.py --json) and bench.py drops the ncu block when it no longer matches.
KERNEL_SOURCES = {
    "inter": ["mptc_inter_wide.cu", "mptc_inter.cu", "mptc_inter_tile.cuh", "mptc_uniform_eval.cuh", "mptc_device.cuh", "mptc_kernels.h"],
    "intra": ["mptc_intra_rows.cu", "mptc_uniform_eval.cuh", "mptc_device.cuh", "mptc_kernels.h"],
}


def kernel_source_sha(stage: str) -> str:
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES[stage]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = ["-DMPTC_PHASE_TIMING=" + os.environ["MPTC_PHASE_TIMING"]] if os.environ.get("MPTC_PHASE_TIMING") else []   # profiling builds only
    extra += os.environ.get("MPTC_EXTRA_NVCC_FLAGS", "").split()
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
