"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

    python pytorch-simclr_b200/build.py [--force] [--verbose]

Output: pytorch-simclr_b200/lib/libsimclr_b200.so  (git-ignored, travels to the GPU box with gpurun).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsimclr_b200.so")
TRACE_LIB_PATH = os.path.join(LIB_DIR, "libsimclr_b200_trace.so")      # the same library with the debug stamps compiled in
STAMP = LIB_PATH + ".stamp"
SOURCES = ["capi.cu"]
DEPS = ["capi.cu", "contrastive_kernels.cuh", "aux_kernels.cuh", "head_kernels.cuh", "selftest.cuh", "probes.cuh", "sm100_ptx.cuh",
        "row_math.cuh",
        os.path.join("..", "..", "include", "simclr_b200.h"), os.path.join("..", "..", "include", "simclr_b200_debug.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if os.path.isfile(cand):
        return cand
    found = shutil.which("nvcc")
    if not found:
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return found


_COMMENT = re.compile(rb"//[^\n]*|/\*.*?\*/", re.S)


def _digest() -> str:
    """Stamp of the library: the compiler flags and the sources with comments and blank space removed, so that a comment
    edit does not orphan the profiles that are keyed to the stamp (profiles/ncu_traffic.json).  A "//" inside a string
    literal is stripped as well -- harmless for a digest."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for dep in DEPS:
        with open(os.path.join(CSRC, dep), "rb") as f:
            h.update(b" ".join(_COMMENT.sub(b"", f.read()).split()))
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(STAMP) and open(STAMP).read().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    # the tracing variant for tools/trace_timeline.py and tools/cta_timeline.py, built alongside
    trace_cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DSIMCLR_TRACE=1", "-o", TRACE_LIB_PATH] + \
                [os.path.join(CSRC, s) for s in SOURCES]
    trace_proc = subprocess.Popen(trace_cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    trace_out, _ = trace_proc.communicate()
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libsimclr_b200.so")
    if trace_proc.returncode != 0:
        sys.stderr.write(trace_out)
        raise RuntimeError("nvcc failed building libsimclr_b200_trace.so")
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(proc.stdout + proc.stderr)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose))
