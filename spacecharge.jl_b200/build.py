"""Builds lib/libspacecharge_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, so the
shared object travels to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libspacecharge_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# translation unit -> extra flags.  The particle and Green kernels follow the reference's
# un-contracted arithmetic (no fused multiply-add), the FFT passes may contract.
UNITS = {
    "api.cu": [],
    "fft_passes_f32.cu": [],
    "fft_passes_f64.cu": [],
    "green.cu": ["-fmad=false"],
    "particles.cu": ["-fmad=false"],
    "sorted.cu": ["-fmad=false"],
}


def _headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "spacecharge_b200.h"))
    return out


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force: bool = False, verbose: bool = False, tag: str = "", defines=()) -> str:
    """tag/defines build an experimental variant next to the product library
    (lib/libspacecharge_b200_<tag>.so), selected at run time with SCB_LIB=<path>."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = OBJDIR + ("_" + tag if tag else "")
    lib_path = LIB if not tag else os.path.join(LIBDIR, "libspacecharge_b200_%s.so" % tag)
    os.makedirs(objdir, exist_ok=True)
    hdrs = _headers()
    defs = ["-D" + d for d in defines]

    def compile_one(item):
        name, extra = item
        src = os.path.join(CSRC, name)
        obj = os.path.join(objdir, name.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + COMMON + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, r.stdout, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(UNITS))) as ex:
        objs = list(ex.map(compile_one, UNITS.items()))
    if force or _stale(lib_path, objs):
        cmd = [NVCC, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib_path


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--tag", default="")
    ap.add_argument("-D", action="append", default=[])
    a = ap.parse_args()
    print(build_library(force=a.force, verbose=a.v, tag=a.tag, defines=a.D))
