"""Time the cold-geometry solve (Green spectrum rebuilt) several times: wall clock and library events."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
scb = load_package()
grid = (256, 256, 256)
mesh = scb.Mesh3D(grid, (-5e-3,) * 3, (5e-3,) * 3)
mesh.rho.normal_()
hd = mesh.handle
hd.enable_timing(True)
scb.solve_(mesh); torch.cuda.synchronize()
for i in range(4):
    t0 = time.perf_counter(); hd.drop_green_cache(); t1 = time.perf_counter()
    scb.solve_(mesh); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    t = hd.timing()
    print("drop %.2f ms, solve call (host) %.2f ms, until sync %.2f ms, events: solve %.2f green %.2f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t1), t["solve_ms"], t["green_ms"]))
for i in range(3):
    t1 = time.perf_counter(); scb.solve_(mesh); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    t = hd.timing()
    print("warm: host %.3f ms, until sync %.3f ms, events solve %.3f" % (1e3 * (t2 - t1), 1e3 * (t3 - t1), t["solve_ms"]))
