"""Tracking-loop step (mesh re-fitted to the bunch every step => Green spectrum rebuilt) for one workload, as bench.py's
cold_geometry.remesh_step_ms measures it.  Run once with SCB_GREEN_OVERLAP=0 and once with =1 to see what building the
spectrum on a second stream during the deposit hides.
usage: python tools/remesh_timing.py [workload] [dtype]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import WORKLOADS, SIGMA, QTOT  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "large"
dt = sys.argv[2] if len(sys.argv) > 2 else "f64"
npart, grid, cath, zshift = WORKLOADS[wl]
scb = load_package()
tdt = torch.float64 if dt == "f64" else torch.float32
gen = torch.Generator(device="cuda")
gen.manual_seed(42)
x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * SIGMA for _ in range(3))
z += zshift * SIGMA
q = torch.full((npart,), QTOT / npart, device="cuda", dtype=tdt)
ex, ey, ez = (torch.empty_like(x) for _ in range(3))
mesh = scb.Mesh3D(grid, x, y, z, T=np.float64 if dt == "f64" else np.float32)
xs = [x * f for f in (1.0, 1.0001, 0.9999, 1.0002, 1.0003, 0.9998)]
for k in range(2):
    mesh.remesh_(xs[k], y, z)
    scb.step_(mesh, xs[k], y, z, q, ex, ey, ez, at_cathode=cath)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for k in range(6):
        mesh.remesh_(xs[k], y, z)
        scb.step_(mesh, xs[k], y, z, q, ex, ey, ez, at_cathode=cath)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 6)
for _ in range(3):
    scb.step_(mesh, xs[5], y, z, q, ex, ey, ez, at_cathode=cath)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    scb.step_(mesh, xs[5], y, z, q, ex, ey, ez, at_cathode=cath)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"workload": wl, "dtype": dt, "overlap": os.environ.get("SCB_GREEN_OVERLAP", "default"),
                  "remesh_step_ms": round(best, 4), "warm_step_ms": round(e0.elapsed_time(e1) / 5, 4)}))
