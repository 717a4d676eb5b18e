"""Python twin of the reference's benchmark/full_pipeline_profile.jl: one full pipeline (deposit + solve +
interpolate) per backend, with the same flags.

    python tools/full_pipeline_profile.py [--cpu | --gpu | --both] [--workload analytic] [--dtype f64]

--gpu  per-stage and per-pass device times of this library (CUDA events inside the C ABI) for one warm and one
       cold-geometry pipeline, plus the launch count; run the same command under
       `ncu --metrics gpu__time_duration.sum --clock-control none` for the per-kernel list (B200_PROFILING recipe).
--cpu  function-level profile (cProfile, like Julia's Profile in the reference script) of the CPU restatement of the
       reference's structure, run through `bench.py --impl reference` (the one place that may execute oracle/).
--both is the default, as in the reference script (benchmark/full_pipeline_profile.jl:57-62)."""
import argparse
import cProfile
import json
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bunch(n, dtype, sigma=1.0e-3, total_charge=1.0e-9):
    rng = np.random.Generator(np.random.PCG64(42))
    x, y, z = (rng.standard_normal(n).astype(dtype) * dtype(sigma) for _ in range(3))
    return x, y, z, np.full(n, total_charge / n, dtype=dtype)


def profile_gpu(grid, n, dtype):
    import torch
    from __graft_entry__ import load_package
    scb = load_package()
    x, y, z, q = (torch.from_numpy(a).cuda() for a in bunch(n, dtype))
    mesh = scb.Mesh3D(grid, x, y, z, T=dtype, total_charge=1.0e-9)
    hd = mesh.handle
    hd.enable_timing(True)
    out = {}
    for label in ("cold_geometry", "warm"):
        l0 = hd.launch_count()
        scb.deposit_(mesh, x, y, z, q)
        scb.solve_(mesh)
        scb.interpolate_field(mesh, x, y, z)
        torch.cuda.synchronize()
        t = hd.timing()
        out[label] = {"deposit_ms": t["deposit_ms"], "solve_ms": t["solve_ms"], "interpolate_ms": t["interpolate_ms"],
                      "green_build_ms": t["green_ms"], "passes_ms": dict(zip(("F1", "F2", "Z", "B2", "B3"), t["pass_ms"])),
                      "launches": hd.launch_count() - l0}
    print(json.dumps({"backend": "b200", "grid": grid, "particles": n, **out}, indent=1))


def profile_cpu(workload, top=14):
    """The oracle is test / baseline infrastructure: it is executed only through bench.py's reference arm, here under
    cProfile (the counterpart of Julia's Profile in the reference script)."""
    import subprocess
    cmd = [sys.executable, "-m", "cProfile", "-s", "cumulative", os.path.join(ROOT, "bench.py"), "--impl", "reference",
           "--workload", workload, "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    lines = r.stdout.splitlines()
    start = next((i for i, ln in enumerate(lines) if "function calls" in ln), 0)
    print("backend: CPU restatement of the reference's structure (bench.py --impl reference --workload %s)" % workload)
    # header, then the rows of the pipeline itself (module import time is noise at the small workloads)
    keep = [ln for ln in lines[start + 5:] if any(k in ln for k in ("cpu_reference.py", "pyduccfft", "cpu_reference_run"))]
    print("\n".join(lines[start:start + 5] + keep[:top]))


def main():
    from bench import WORKLOADS
    ap = argparse.ArgumentParser()
    g = ap.add_mutually_exclusive_group()
    g.add_argument("--cpu", action="store_true")
    g.add_argument("--gpu", action="store_true")
    g.add_argument("--both", action="store_true")
    ap.add_argument("--workload", default="analytic", choices=sorted(WORKLOADS), help="bench.py workload (BASELINE.json configs)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    a = ap.parse_args()
    n, grid = WORKLOADS[a.workload][:2]
    dtype = np.float64 if a.dtype == "f64" else np.float32
    if a.cpu or a.both or not a.gpu:
        profile_cpu(a.workload)
    if a.gpu or a.both or not a.cpu:
        profile_gpu(grid, n, dtype)


if __name__ == "__main__":
    main()
