"""Times scb_sort_particles and scb_permute on the config-5 bunch (random order, then the already ordered bunch) for the
CURRENT library build / environment; one JSON line.  usage: python tools/sort_probe.py [f64|f32] [label]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
label = sys.argv[2] if len(sys.argv) > 2 else ""
npart = int(float(os.environ.get("SCB_PROBE_NP", "1e8")))
scb = load_package()
tdt = torch.float64 if dt == "f64" else torch.float32
gen = torch.Generator(device="cuda")
gen.manual_seed(42)
x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * 1e-3 for _ in range(3))
q = torch.full((npart,), 1e-9 / npart, device="cuda", dtype=tdt)
mesh = scb.Mesh3D((256, 256, 256), x, y, z, T=np.float64 if dt == "f64" else np.float32)


def timed(fn, reps=4):
    best, out = None, None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return round(best, 4), out


rep = {"label": label, "dtype": dt, "lib": os.path.basename(os.environ.get("SCB_LIB", "product"))}
rep["sort_random"], perm = timed(lambda: scb.sort_particles(mesh, x, y, z))
rep["permute4_random"], (sx, sy, sz, sq) = timed(lambda: scb.permute(perm, x, y, z, q, handle=mesh.handle))
rep["sort_ordered"], perm2 = timed(lambda: scb.sort_particles(mesh, sx, sy, sz))
rep["permute4_ordered"], _ = timed(lambda: scb.permute(perm2, sx, sy, sz, sq, handle=mesh.handle))
# a drifted bunch: the steady state of a tracking loop that re-sorts every few steps
d = [float(v) for v in mesh.delta]
mx, my, mz = (s + torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * (0.1 * dd) for s, dd in zip((sx, sy, sz), d))
for a, lo, hi in zip((mx, my, mz), mesh.min_bounds, mesh.max_bounds):
    a.clamp_(float(lo), float(hi))
rep["sort_drift0.1"], perm3 = timed(lambda: scb.sort_particles(mesh, mx, my, mz))
rep["permute4_drift0.1"], _ = timed(lambda: scb.permute(perm3, mx, my, mz, sq, handle=mesh.handle))
ok = bool((perm2.long() == torch.arange(npart, device="cuda")).all())
rep["resort_is_identity"] = ok
print(json.dumps(rep))
