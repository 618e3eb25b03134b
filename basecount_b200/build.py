"""Build the CUDA C-ABI library in-tree:  python -m basecount_b200.build

nvcc cross-compiles for sm_100a without a GPU; the resulting
basecount_b200/csrc/libbasecount_b200.so is git-ignored but travels with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libbasecount_b200.so")
SOURCES = ["bc_api.cu"]
HEADERS = ["bc_common.cuh", "k1_count.cuh", "k1_fast.cuh", "k2_stats.cuh", "k3_reduce.cuh", "bam_decode.h", "cigar_canon.h", "nccl_dyn.h", "bam_index.h", "inflate_fast.h", "tsv_format.h", 
           os.path.join("..", "..", "include", "basecount_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # K2 must not contract a*b+c: float64 results follow the reference's operation order
    "-Xcompiler", "-fPIC,-O3,-pthread",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """`out` / `defines` build an experimental variant next to the product library (see BASECOUNT_B200_LIB)."""
    if out == OUT and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES + ["-lz", "-ldl"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _out = next((a[6:] for a in sys.argv[1:] if a.startswith("--out=")), OUT)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out, defines=_defs))
