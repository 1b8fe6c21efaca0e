"""Builds libacg.so (CUDA kernels + C ABI + host-side C++ mirror) in-tree for sm_100a.

    python arithmetic-circuits_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repository snapshot.  cudart is linked statically so the library loads on a CPU-only machine
(compute calls then fail with ACG_ERR_NO_DEVICE -- there is no CPU fallback)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libacg.so")
OBJ_DIR = os.path.join(HERE, "_build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["r1cs_kernels.cu", "ntt_kernels.cu", "lagrange_kernels.cu", "witness_kernels.cu", "poly_kernels.cu", "linear_kernels.cu", "abi.cu"]
CPP_SOURCES = ["host/circuit.cpp", "host/synth.cpp"]
HEADERS = ["fr.cuh", "fr_constants.inc", "dev.cuh", "kernels.h", "host/fr_host.hpp", "host/circuit.hpp",
           "../../include/acg.h"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"] + os.environ.get("ACG_NVCC_EXTRA", "").split()


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    procs = []
    for src in CU_SOURCES + CPP_SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _newer(obj, [sp] + hdrs):
            cmd = [NVCC] + ARCH + COMMON + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", sp, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libacg build failed")
    if force or procs or _newer(OUT, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static", "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
