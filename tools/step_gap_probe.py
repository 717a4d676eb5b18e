"""Where do the ~0.2 ms between the fused step and the sum of its stage times go?  Times N back-to-back scb_step calls
with CUDA events (N = 10 and 50), the same steps replayed from a CUDA graph, and the stage sequence with the library's
own events.  usage: python tools/step_gap_probe.py [f64|f32]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import WORKLOADS, SIGMA, QTOT  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
npart, grid, cath, zshift = WORKLOADS["large"]
scb = load_package()
tdt = torch.float64 if dt == "f64" else torch.float32
gen = torch.Generator(device="cuda")
gen.manual_seed(42)
x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * SIGMA for _ in range(3))
q = torch.full((npart,), QTOT / npart, device="cuda", dtype=tdt)
ex, ey, ez = (torch.empty_like(x) for _ in range(3))
mesh = scb.Mesh3D(grid, x, y, z, T=np.float64 if dt == "f64" else np.float32)
out = {"dtype": dt}


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


step = lambda: scb.step_(mesh, x, y, z, q, ex, ey, ez)
for _ in range(3):
    step()
out["step_x10_ms"] = [round(timed(step, 10), 4) for _ in range(3)]
out["step_x50_ms"] = round(timed(step, 50), 4)
hd = mesh.handle
hd.enable_timing(True)
best = None
for _ in range(10):
    step()
    t = hd.timing()
    cur = (t["deposit_ms"], t["solve_ms"], t["interpolate_ms"])
    best = cur if best is None else tuple(min(a, b) for a, b in zip(best, cur))
hd.enable_timing(False)
out["stages_min_ms"] = [round(v, 4) for v in best]
out["stages_sum_ms"] = round(sum(best), 4)
# the same ten steps from a CUDA graph (no launch gaps, no host work between kernels)
try:
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(10):
                step()
    torch.cuda.synchronize()
    g.replay()
    out["graph_x10_ms"] = [round(timed(g.replay, 1) / 10, 4) for _ in range(3)]
except Exception as exc:   # capture may be refused (a library call that is illegal during capture)
    out["graph_error"] = repr(exc)[:300]
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "step_gap_%s.json" % dt), "w") as f:
    json.dump(out, f)
