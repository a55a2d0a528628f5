"""Compile the UNCHANGED kernel sources with g++ against tests/emu/cuda_emu.h.

DEBUG HARNESS, not a backend: the result (tests/emu/_build/libinsr_emu.so) exposes the same
C ABI on host pointers so that tests can exercise kernel indexing / layouts and the Python
host logic in the GPU-less build container.  The product package never loads it."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "insr_pde_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libinsr_emu.so")


def build_emu(force: bool = False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
            os.path.join(ROOT, "include", "insr_b200.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-DINSR_CPU_EMU", "-DINSR_SINGLE_TU",
           "-Wno-unknown-pragmas", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           "-x", "c++", os.path.join(CSRC, "insr_abi.cu"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulation library")
    return LIB


if __name__ == "__main__":
    print(build_emu(force=True))
