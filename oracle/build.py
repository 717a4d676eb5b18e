"""Builds oracle/_port/libref_port.so (gcc -O2 -fopenmp, no fast-math, no FMA contraction).

`oracle/_ref/` is reserved for a build of the reference's own sources; SpaceCharge.jl is pure Julia
and Julia is not installed here, so nothing can be built there (stated in DESIGN.md)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_port")
LIB = os.path.join(OUT, "libref_port.so")
SRC = os.path.join(HERE, "ref_port.c")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
