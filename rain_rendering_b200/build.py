"""Builds librain_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librain_b200.so")
SOURCES = ["rr_api.cu", "rr_kernels.cu", "rr_png_gpu.cu", "rr_sim.cu", "rr_host.cpp", "rr_host_xml.cpp", "rr_host_png.cpp"]
HEADERS = ["rr_types.h", "rr_cvmath.h", "rr_streak_geom.h", "rr_kernels.cuh", "rr_png_gpu.cuh", "rr_host_deflate.h", "rr_host_inflate.h", "rr_viridis.h",
           os.path.join("..", "..", "include", "rain_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",            # keep the reference's float64 operation order (no FMA contraction)
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "static"]


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built and there is no CPU fallback")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    # Several processes may get here at once (one rank per GPU under torchrun, the launcher's children): one of them builds,
    # into a temporary file that is renamed over the library only when it is complete; the others wait on the lock and then
    # find a fresh library.  A half-written .so is never visible under the final name.
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lz"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                print(res.stdout + res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
