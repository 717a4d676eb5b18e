"""A tracking loop on the BASELINE config-5 bunch: what keeping the bunch ordered by cell is worth over a run of steps.

Every step = deposit + solve + interpolate (scb_step) + a ballistic drift x += v (the caller's push; velocities are drawn
once so that a step moves a particle by N(0, (d * cell)^2) per axis).  Compared, for several per-step drifts d and
re-sort intervals K, as average milliseconds per step in the STEADY STATE, including the re-sorts (seven arrays are
permuted: x, y, z, q and the three velocities) and the drift kernels (three torch adds, 1.2 ms per step):

  random   the bunch stays in its random order, default kernels (nothing to maintain)
  ordered  scb_sort_particles + scb_permute of x, y, z, q, vx, vy, vz every K steps, SCB_ORDER_CELL kernels in between

usage: python tools/tracking_loop_demo.py [f64|f32] [steps]      (writes gpurun_out/tracking_loop_demo_<dtype>.json)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    npart = int(float(os.environ.get("SCB_PROBE_NP", "1e8")))
    grid = (256, 256, 256)
    scb = load_package()
    tdt = torch.float64 if dt == "f64" else torch.float32
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    x0, y0, z0 = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * 1e-3 for _ in range(3))
    q = torch.full((npart,), 1e-9 / npart, device="cuda", dtype=tdt)
    # fixed geometry with head-room for the drift (a real loop would call remesh_; kept out to isolate the ordering)
    lo = tuple(1.3 * float(a.min()) for a in (x0, y0, z0))
    hi = tuple(1.3 * float(a.max()) for a in (x0, y0, z0))
    mesh = scb.Mesh3D(grid, lo, hi, T=np.float64 if dt == "f64" else np.float32)
    cell = [float(v) for v in mesh.delta]
    outs = [torch.empty_like(x0) for _ in range(3)]
    unit = [torch.randn(npart, generator=gen, device="cuda", dtype=tdt) for _ in range(3)]
    rep = {"dtype": dt, "particles": npart, "grid": list(grid), "steps": steps, "runs": []}

    def run(d, K):
        """average ms per step; K = 0: random order, default kernels"""
        x, y, z = x0.clone(), y0.clone(), z0.clone()
        v = [u * (d * c) for u, c in zip(unit, cell)]
        qq = q.clone()
        scb.set_particle_order(mesh, "cell" if K else "random")
        perm = torch.empty(npart, dtype=torch.int32, device="cuda")
        spare = [torch.empty_like(x) for _ in range(7)]
        def resort():
            nonlocal x, y, z, qq, v, spare
            scb.sort_particles(mesh, x, y, z, out=perm)
            new = scb.permute(perm, x, y, z, qq, *v, handle=mesh.handle, out=spare)
            spare = [x, y, z, qq, *v]
            x, y, z, qq, v = new[0], new[1], new[2], new[3], list(new[4:])

        def step():
            scb.step_(mesh, x, y, z, qq, *outs)
            x.add_(v[0]); y.add_(v[1]); z.add_(v[2])

        try:
            # steady state: the timed window starts with a bunch that was ordered K steps ago (the first sort of the
            # random bunch is a one-off and stays outside)
            if K:
                resort()
                for _ in range(K):
                    step()
            else:
                step()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for s in range(steps):
                if K and s % K == 0:
                    resort()
                step()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / steps
        finally:
            scb.set_particle_order(mesh, "random")

    for d in (0.002, 0.01, 0.03):
        row = {"drift_cells_per_step": d, "random_order_ms": round(run(d, 0), 3), "ordered_ms_by_resort_interval": {}}
        for K in (1, 2, 4, 8, 12, 24):
            row["ordered_ms_by_resort_interval"][str(K)] = round(run(d, K), 3)
        rep["runs"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tracking_loop_demo_%s.json" % dt), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
