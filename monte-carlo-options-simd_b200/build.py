"""In-tree build of libmcb200.so for sm_100a (nvcc cross-compiles without a GPU).

    python monte-carlo-options-simd_b200/build.py [--force] [--verbose]

The library lands in monte-carlo-options-simd_b200/lib/libmcb200.so: git-ignored, but it travels to the GPU box with
the gpurun snapshot, so the box never needs to compile.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, os.environ.get("MCB200_LIB_NAME", "libmcb200.so"))     # alternate name: tuning builds only
SOURCES = ["mcb200_api.cu", "mcb200_multi.cu", "mcb200_microbench.cu", "mcb200_plan.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"   # the image's CC/CXX wrappers lack libgomp etc.


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmcb200.so cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    if not (force or stale()):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("MCB200_NVCC_FLAGS", "").split()          # tuning experiments only (e.g. -DMCB_QUAD_UNROLL=4)
    cmd = [nvcc(), "-ccbin", HOST_CXX, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall", *extra,
           "-shared", "-cudart", "static", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
