"""Minimal driver for ncu: a few full steps of one workload through the C ABI (no timing, no CPU work).
usage: python tools/profile_step.py [workload] [dtype] [steps]"""
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
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
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
for _ in range(steps):
    scb.step_(mesh, x, y, z, q, ex, ey, ez, at_cathode=cath)
torch.cuda.synchronize()
print("done", mesh.handle.launch_count())
