"""Builds baseline/naive_gpu/libref_structure.so (nvcc + cuFFT).  Bench/test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libref_structure.so")
SRC = os.path.join(HERE, "ref_structure.cu")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC, "-lcufft"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
