"""Builds libhtsp_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.
No torch dependency: the library only uses the CUDA runtime (statically linked)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libhtsp_b200.so")
STAMP = os.path.join(PKG_DIR, ".libhtsp_b200.stamp")
SOURCES = ["abi.cu", "extract.cu", "encoder_f32.cu", "decode_prep.cu", "decode_f32.cu",
           "decode_mma.cu", "decode_tc.cu", "decode_tc2.cu", "select.cu", "env_state.cu", "env_fragment.cu", "upper_policy.cu", "gemm_tc.cu", "encoder_tc.cu", "enc_attention_mma.cu", "bn_train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
              "-cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/htsp_b200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if (not force and os.path.isfile(LIB_PATH) and os.path.isfile(STAMP)
            and open(STAMP).read().strip() == dig):
        return LIB_PATH
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    os.makedirs(os.path.join(PKG_DIR, "build"), exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(PKG_DIR, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-Xptxas", "-v" if verbose else "-warn-spills",
               "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:   # one nvcc per translation unit
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [_nvcc(), "-shared", *flags, "-o", LIB_PATH, *objs]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
