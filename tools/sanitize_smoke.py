"""Small end-to-end exercise for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family
once, small sizes.  usage: compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

scb = load_package()
rng = np.random.default_rng(0)
for T in (np.float64, np.float32):
    for grid, n, cath in (((12, 9, 20), 5000, True), ((32, 32, 32), 40000, False), ((5, 70, 6), 300, False)):
        x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
        z = z + 6e-3
        q = np.full(n, 1e-9 / n)
        d = [torch.from_numpy(a.astype(T)).cuda() for a in (x, y, z, q)]
        mesh = scb.Mesh3D(grid, *d[:3], T=T, gamma=2.0)
        for mode in (1, 2, 3):
            mesh.handle.lib.scb_drop_green_cache(mesh.handle.h)
            scb.deposit_(mesh, *d)
            scb.solve_(mesh, at_cathode=cath)
            out = scb.interpolate_field(mesh, *d[:3])
        scb.solve_freespace_(mesh, (1e-4, 0.0, 2e-4))
        outs = [torch.empty_like(d[0]) for _ in range(3)]
        scb.step_(mesh, *d, *outs)
        h = [np.empty(n, dtype=T) for _ in range(3)]
        scb.step_host_(mesh, *(a.astype(T) for a in (x, y, z, q)), *h)
        g = scb.get_green_function_((2 * grid[0], 2 * grid[1], 2 * grid[2]), mesh.delta, 2.0, 2, (0, 0, 1e-3), T=T)
        torch.cuda.synchronize()
        assert np.isfinite(h[0]).all() and bool(torch.isfinite(mesh.efield).all())
print("sanitize smoke ok")
