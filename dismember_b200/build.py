"""Build recipe for libdismember_gpu.so (sm_100a only, in-tree so it travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdismember_gpu.so")
SOURCES = ["capi.cu", "dr.cu", "dr_train.cu", "train.cu", "shard.cu", "otm_deepfm.cu", "cluster.cu"]
OBJ = os.path.join(HERE, "..", "build", "obj")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",                      # only explicit fma intrinsics fuse (dmg_math.cuh)
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]
LINK_FLAGS = ["-shared", "-ldl"]


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
    # one nvcc -c per translation unit, in parallel; an object is reused when neither its source, a header of csrc/ nor the
    # public header is newer (objects of a build with extra flags live in their own directory)
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(OBJ, "default" if not extra else "x" + str(abs(hash(" ".join(extra)))))
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [os.path.join(HERE, "..", "include", "dismember_gpu.h"), __file__]
    t_hdr = max(os.path.getmtime(f) for f in hdrs)
    log = []

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(t_hdr, os.path.getmtime(path)):
            return obj, 0
        res = subprocess.run([nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(res.stdout)
        return obj, res.returncode
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        done = list(ex.map(compile_one, SOURCES))
    if verbose or any(rc for _, rc in done):
        print("".join(log))
    if any(rc for _, rc in done):
        raise RuntimeError("nvcc failed building libdismember_gpu.so")
    res = subprocess.run([nvcc()] + NVCC_FLAGS + LINK_FLAGS + ["-o", out] + [o for o, _ in done], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode:
        print(res.stdout)
        raise RuntimeError("nvcc failed linking libdismember_gpu.so")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
