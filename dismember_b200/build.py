"""Build recipe for libdismember_gpu.so (sm_100a only, in-tree so it travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdismember_gpu.so")
SOURCES = ["capi.cu", "dr.cu", "train.cu", "shard.cu", "otm_deepfm.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",                      # only explicit fma intrinsics fuse (dmg_math.cuh)
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-ldl",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dismember_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = LIB) -> str:
    if out == LIB and not force and not stale():
        return LIB
    extra = os.environ.get("DMG_NVCC_EXTRA", "").split()        # e.g. -DDMG_FAST_TIMING for the phase timers
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        print(res.stdout)
    if res.returncode:
        raise RuntimeError("nvcc failed building libdismember_gpu.so")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
