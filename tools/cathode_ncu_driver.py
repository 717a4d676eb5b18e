"""Short driver for ncu captures of the cathode solve (BASELINE config 4: 1e7 particles, 128 x 128 x 256, image-charge
term): `steps` steps with at_cathode=True.  usage: python tools/cathode_ncu_driver.py [f64|f32] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scb = load_package()
tdt = torch.float64 if dt == "f64" else torch.float32
gen = torch.Generator(device="cuda")
gen.manual_seed(7)
npart = 10_000_000
x, y = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * 1e-3 for _ in range(2))
z = torch.rand(npart, generator=gen, device="cuda", dtype=tdt) * 2e-3 + 1e-4      # in front of the cathode plane z = 0
q = torch.full((npart,), 1e-9 / npart, device="cuda", dtype=tdt)
mesh = scb.Mesh3D((128, 128, 256), x, y, z, T=np.float64 if dt == "f64" else np.float32)
outs = [torch.empty_like(x) for _ in range(3)]
for _ in range(steps):
    scb.step_(mesh, x, y, z, q, *outs, at_cathode=True)
torch.cuda.synchronize()
print("done")
