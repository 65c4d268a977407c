"""Build the C-ABI CUDA library in-tree:  passt_b200/lib/libpasst_b200.so  (sm_100a only).

The library has no torch / python dependency: it is plain nvcc output exposing the ``extern "C"`` entry points
declared in include/passt_b200.h.  The built .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpasst_b200.so")
SOURCES = ["gemm.cu", "gemm2.cu", "mel.cu", "rowops.cu", "attn_fwd.cu", "attn_fwd2.cu", "attn_fwd3.cu", "attn_bwd.cu", "optim.cu", "loss.cu", "waveaug.cu", "fp32tier.cu", "fp32tier_bwd.cu", "patch_embed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
]
# --use_fast_math would change logf/erff/div accuracy in the parity-critical kernels; keep IEEE there.
PRECISE = {"mel.cu", "rowops.cu", "gemm.cu", "gemm2.cu", "attn_fwd.cu", "attn_fwd2.cu", "attn_fwd3.cu", "attn_bwd.cu", "optim.cu", "loss.cu", "waveaug.cu", "fp32tier.cu", "fp32tier_bwd.cu", "patch_embed.cu"}


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    return "nvcc"


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest(deps)
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        flags = [f for f in NVCC_FLAGS if not (f == "--use_fast_math" and src in PRECISE)]
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(LIBDIR, src + ".ptxas.log"), "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
