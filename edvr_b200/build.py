"""Build libedvr_b200.so in-tree with nvcc for sm_100a (no torch headers involved)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libedvr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
         "-ccbin", "/usr/bin/g++"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "..", "include", "edvr_b200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB, os.path.join(CSRC, "capi.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libedvr_b200.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
