"""Short driver for ncu captures of the cell-ordered particle kernels (and the sort) at BASELINE config 5:
builds the ordered bunch, then runs `steps` steps with SCB_ORDER_CELL.  usage: python tools/sorted_ncu_driver.py [f64|f32] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
npart = int(float(os.environ.get("SCB_PROBE_NP", "1e8")))
scb = load_package()
tdt = torch.float64 if dt == "f64" else torch.float32
gen = torch.Generator(device="cuda")
gen.manual_seed(42)
x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * 1e-3 for _ in range(3))
q = torch.full((npart,), 1e-9 / npart, device="cuda", dtype=tdt)
mesh = scb.Mesh3D((256, 256, 256), x, y, z, T=np.float64 if dt == "f64" else np.float32)
perm, sx, sy, sz, sq = scb.sort_particles_(mesh, x, y, z, q)
del x, y, z, q
outs = [torch.empty_like(sx) for _ in range(3)]
scb.set_particle_order(mesh, "cell")
for _ in range(steps):
    scb.step_(mesh, sx, sy, sz, sq, *outs)
torch.cuda.synchronize()
print("done")
