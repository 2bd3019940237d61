"""Build the CUDA library ``brille_b200/libbrille_b200.so`` for sm_100a with nvcc (in-tree, no torch)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbrille_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# per-file extra flags: the decision arithmetic of the locate stage must not be contracted into FMAs
SOURCES = {
    "locate.cu": ["-fmad=false"],
    "interp.cu": [],
    "cellinterp.cu": [],
    "cellinterp_tma.cu": [],
    "consumer.cu": [],
    "sortpairs.cu": ["-fmad=false"],  # cost matrices in the reference's rounding (x86-64 baseline: no fused multiply-add)
    "capi.cu": [],
}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "brille_b200.h"))
    objs = []
    rebuilt = False
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc(), *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            rebuilt = True
        objs.append(o)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
