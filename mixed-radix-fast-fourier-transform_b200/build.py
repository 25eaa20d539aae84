"""Build libhsfft_b200.so in-tree: nvcc (sm_100a) for the CUDA side, gcc for the plain-C host side.

    python mixed-radix-fast-fourier-transform_b200/build.py [--force] [--ptxas]

The shared library is the drop-in boundary (extern "C", see include/*.h).  cudart is linked
statically, so the only run-time dependency is the driver's libcuda.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HSFFT_VARIANT=<name> HSFFT_DEFS="-DX=1 ..." builds an experiment into libhsfft_b200_<name>.so (A/B runs: HSFFT_LIB=<path>)
VARIANT = os.environ.get("HSFFT_VARIANT", "")
OUT = os.path.join(HERE, "libhsfft_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
BUILD = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-ccbin", GXX] + os.environ.get("HSFFT_DEFS", "").split()
C_FLAGS = ["-O2", "-std=gnu11", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter"]

CU_SOURCES = ["hs_device.cu"]
# the persistent pipeline kernels: one object per two-pass size 2^14..2^20 (same source, -DHS_PIPE_ID), built in parallel
PIPE_SIZES = [14, 15, 16, 17, 18, 19, 20]
C_SOURCES = ["hsfft_plan.c", "hsfft_real.c", "hsfft_convolve.c", "hsfft_dist.c"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_nvcc=()):
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in ("highspeedFFT.h", "real.h", "hsfft_b200.h")]
    jobs = []  # (object, command)
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src + ".o")
        jobs.append((o, [s], [NVCC] + NVCC_FLAGS + list(extra_nvcc) + ["-c", s, "-o", o]))
    for k in PIPE_SIZES:
        s = os.path.join(CSRC, "hs_pipe_inst.cu")
        o = os.path.join(BUILD, "hs_pipe_%d.o" % k)
        jobs.append((o, [s], [NVCC] + NVCC_FLAGS + list(extra_nvcc) + ["-DHS_PIPE_ID=CT_T%d" % k, "-DHS_PIPE_NAME=%d" % k, "-c", s, "-o", o]))
    for src in C_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src + ".o")
        jobs.append((o, [s], [GCC] + C_FLAGS + ["-c", s, "-o", o]))

    def run(job):
        o, deps, cmd = job
        if force or _newer(o, deps + headers):
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
        return o

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(run, jobs))
    if force or _newer(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ccbin", GXX, "-Xlinker", "-Bsymbolic", "-lpthread", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True,
                 extra_nvcc=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ())
    print("built", path)
