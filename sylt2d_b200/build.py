"""Builds libsylt2d_b200.so in-tree with nvcc for sm_100a (B200).  -fmad=false / -ffp-contract=off: the reference's
f32 arithmetic never fuses a*b+c, and bit-exact parity with the CPU oracle depends on that."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsylt2d_b200.so")
SOURCES = ["world.cu", "batch.cu"]
HEADERS = ["common.h", "narrow.cuh", "solver.cuh", "sylt_math.cuh", "sylt_sincos.h", os.path.join("..", "..", "include", "sylt2d_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas", "-shared"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
