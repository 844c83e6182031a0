"""Builds libflic_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python finalproject-losslessimagecompression_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library links the CUDA runtime statically and has no
dependency on torch or Python, so the same file serves ctypes, cgo, JNI or a C program.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libflic_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["cdf_tables.cu", "rans_encode.cu", "rans_decode.cu", "rans_decode_coop.cu", "pack_words.cu", "couple_round.cu",
           "flow_index.cu", "logistic_prob.cu", "capi.cu"]
HEADERS = ["flic_core.cuh", "flic_device.cuh", "flic_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # belt and braces: every rounding-critical op is an explicit intrinsic
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, "flic_b200.h"), __file__]
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for src in sources:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", CSRC, "-I", INCLUDE, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
