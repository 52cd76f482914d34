#!/usr/bin/env python3
"""Builds libkmc_b200.so (CUDA kernels + C ABI) for sm_100a, in-tree.

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false ...

-fmad=false: the parity-critical arithmetic uses explicit __dadd_rn/__dmul_rn already; the flag keeps the
remaining (non-critical) expressions free of contraction as well so a replay never depends on ptxas choices.
Host code is compiled without -march and with -ffp-contract=off (derived constants must equal the reference's).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("KMC_LIB_OUT") or os.path.join(HERE, "libkmc_b200.so")
SRC = [os.path.join(HERE, "csrc", "kmc_engine.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in ("kmc_engine.cu", "kmc_kernels.cu", "kmc_small.cu", "kmc_strips.cu", "kmc_init.cu", "kmc_device.cuh", "kmc_geom.cuh", "kmc_philox.cuh")] + \
       [os.path.join(HERE, "..", "include", "kmc_b200.h")]


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-o", LIB] + SRC
    for k in ("TS", "TTHREADS", "TCAP", "NSURV", "TMINB", "KMC_TILE_TIMING", "CTHREADS", "CSURV", "CMINB", "PMINB", "PTHREADS", "RECMINB", "LIGMINB", "PE_CHUNK", "RP_CHUNK", "RPTHREADS", "SMALL_T", "SMALL_MINB", "SMALL_TIMING", "SMALL_DMAX", "SMALL_SPEC_MAX", "RECDYN", "CX_G", "CX_GROUPS"):          # tile-kernel tuning knobs (experiments only)
        if os.environ.get("KMC_" + k):
            cmd.insert(1, "-D%s=%s" % (k, os.environ["KMC_" + k]))
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
