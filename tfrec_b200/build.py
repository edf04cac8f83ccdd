"""Builds tfrec_b200/libtfrb200.so (CUDA kernels + C ABI) with nvcc for sm_100a, in tree.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtfrb200.so")
SOURCES = ["frontend.cu", "frontend_tc.cu", "frontend_screen.cu", "backend.cu", "backend2.cu", "decim.cu", "decim_fused.cu", "tfr_api.cu"]
HEADERS = ["tfr_dev.h", "demod_dev.cuh", "fir_taps.h", "frontend_common.cuh", "fir_exact.cuh", os.path.join(ROOT, "include", "tfr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "static"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libtfrb200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


CLI = os.path.join(HERE, "tfrec_b200_cli")
HOST_SOURCES = ["main.cpp", "engine.cpp", "fm_demod.cpp", "decoder.cpp"]


def build_cli(force=False):
    """the tfrec-compatible command line: C++ host mirror of the reference's engine/decoder surface over the C ABI"""
    hdir = os.path.join(HERE, "host")
    srcs = [os.path.join(hdir, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(hdir, h) for h in os.listdir(hdir) if h.endswith(".h")] + [LIB]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    cmd = ["g++", "-O2", "-std=c++11", "-Wall", "-o", CLI, *srcs, "-L" + HERE, "-ltfrb200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building tfrec_b200_cli")
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli(force="--force" in sys.argv))
