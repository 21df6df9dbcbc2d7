"""In-tree builds: libcuboid_cuda.so (nvcc, sm_100a) and libcuboid_synth.so (gcc, host-only input generator)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
CUDA_LIB = os.environ.get("CUBOID_CUDA_LIB") or os.path.join(HERE, "libcuboid_cuda.so")   # the override is for developer builds (tools/)
SYNTH_LIB = os.path.join(HERE, "libcuboid_synth.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exactness: no FMA contraction anywhere; IEEE div/sqrt; no flush-to-zero
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))


def build_cuda(force=False, verbose=False):
    srcs = cuda_sources() + [os.path.join(ROOT, "include", "cuboid_cuda.h")]
    if not force and _newer(CUDA_LIB, srcs):
        return CUDA_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC,
                                    "-o", CUDA_LIB, os.path.join(CSRC, "cuboid_cuda.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for d in os.environ.get("CUBOID_NVCC_DEFINES", "").split():   # developer builds, e.g. CUBOID_NVCC_DEFINES=-DCUBOID_ICP_STATS
        cmd.insert(1, d)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return CUDA_LIB


def build_synth(force=False):
    src = os.path.join(CSRC, "synth.c")
    if not force and _newer(SYNTH_LIB, [src]):
        return SYNTH_LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-o", SYNTH_LIB, src, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stdout + r.stderr)
    return SYNTH_LIB


def build_all(force=False):
    build_synth(force)
    build_cuda(force)
