"""Build libinsr_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so sits next to
this file so that it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libinsr_b200.so")
FUSED_SHAPES = [(1, 1), (2, 1), (2, 2)]          # (D, O) pairs of the fused family, one object each
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", *(["-DINSR_TC_DEBUG_FLUSH"] if os.environ.get("INSR_BUILD_DEBUG_FLUSH") else []),
]
OBJ_DIR = os.path.join(HERE, "build")


def _newest_source_mtime():
    m = os.path.getmtime(os.path.abspath(__file__))
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build_library(force: bool = False, verbose: bool = False, extra_flags=()):
    """compile csrc/*.cu -> libinsr_b200.so (objects built in parallel); returns the path.
    Skips when the library is newer than every source."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libinsr_b200.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    jobs = [("insr_abi.o", [os.path.join(CSRC, "insr_abi.cu")])]
    for d, o in FUSED_SHAPES:
        jobs.append((f"siren_fused_{d}{o}.o", [f"-DINSR_INST_D={d}", f"-DINSR_INST_O={o}",
                                               os.path.join(CSRC, "siren_fused_inst.cu")]))
    jobs.append(("siren_tiled.o", [os.path.join(CSRC, "siren_tiled_inst.cu")]))
    jobs.append(("siren_mid.o", [os.path.join(CSRC, "siren_mid_inst.cu")]))
    procs = []
    for obj, args in jobs:
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, *inc, "-c", *args, "-o", os.path.join(OBJ_DIR, obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libinsr_b200.so")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
            *[os.path.join(OBJ_DIR, obj) for obj, _ in jobs], "-o", LIB]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True,
                        extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ()))
