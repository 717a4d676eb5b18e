"""Times deposit / interpolate of the CURRENT library build (SCB_LIB) and environment on the config-5 bunch in cell
order (SCB_ORDER_CELL kernels) -- one line of JSON, for A/B runs of build variants.  usage: python tools/variant_probe.py [f64|f32] [label]"""
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
perm, sx, sy, sz, sq = scb.sort_particles_(mesh, x, y, z, q)
del x, y, z, q, perm
outs = [torch.empty_like(sx) for _ in range(3)]
scb.set_particle_order(mesh, "cell")
hd = mesh.handle
for _ in range(2):
    scb.step_(mesh, sx, sy, sz, sq, *outs)
hd.enable_timing(True)
best = None
for _ in range(5):
    scb.step_(mesh, sx, sy, sz, sq, *outs)
    t = hd.timing()
    cur = {"deposit": t["deposit_ms"], "interpolate": t["interpolate_ms"]}
    best = cur if best is None else {k: min(best[k], cur[k]) for k in cur}
print(json.dumps({"label": label, "dtype": dt, "lib": os.path.basename(os.environ.get("SCB_LIB", "product")),
                  **{k: round(v, 4) for k, v in best.items()}, "checksum": float(outs[0].double().sum())}))
