"""What would a caller gain by keeping its bunch sorted by cell?  (first item.)

The particle passes run at the ceiling of scattered 32-byte sector traffic for particles in RANDOM order; a tracking loop
can keep its bunch approximately cell-sorted (re-sorting every few turns).  This probe times the EXISTING kernels on
the BASELINE config-5 bunch in four orders -- random (the benchmark's), fully cell-sorted, sorted by z plane only, and
cell-sorted then shuffled inside windows of W particles (a stand-in for an order that has degraded since the last
sort) -- plus the cost of the sort itself (torch.sort of the linear cell index and the gathers of x, y, z, q), so that the
break-even number of steps per sort can be read off.  Results must not depend on the order: the interpolated field is
compared with the random-order run through the permutation.

usage: python tools/sorted_input_probe.py [f64|f32] [particles]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import WORKLOADS, SIGMA, QTOT  # noqa: E402


def stage_times(scb, mesh, x, y, z, q, outs, reps=5):
    hd = mesh.handle
    for _ in range(2):
        scb.step_(mesh, x, y, z, q, *outs)
    hd.enable_timing(True)
    best = None
    for _ in range(reps):
        scb.step_(mesh, x, y, z, q, *outs)
        t = hd.timing()
        cur = {"deposit": t["deposit_ms"], "solve": t["solve_ms"], "interpolate": t["interpolate_ms"]}
        best = cur if best is None else {k: min(best[k], cur[k]) for k in cur}
    hd.enable_timing(False)
    best["step"] = sum(best.values())
    return {k: round(v, 4) for k, v in best.items()}


def main():
    dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
    npart, grid, _, _ = WORKLOADS["large"]
    if len(sys.argv) > 2:
        npart = int(sys.argv[2])
    scb = load_package()
    tdt = torch.float64 if dt == "f64" else torch.float32
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * SIGMA for _ in range(3))
    q = torch.full((npart,), QTOT / npart, device="cuda", dtype=tdt)
    mesh = scb.Mesh3D(grid, x, y, z, T=np.float64 if dt == "f64" else np.float32)
    outs = [torch.empty_like(x) for _ in range(3)]
    report = {"dtype": dt, "particles": npart, "grid": list(grid)}
    report["random"] = stage_times(scb, mesh, x, y, z, q, outs)
    ref = [o.clone() for o in outs]

    ix, iy, iz = scb.cell_indices(mesh, x, y, z)
    key = (ix + grid[0] * (iy + grid[1] * iz)).to(torch.int32)
    del ix, iy
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    perm = torch.sort(key).indices
    ev[1].record()
    sx, sy, sz, sq = (a[perm] for a in (x, y, z, q))
    ev[2].record()
    torch.cuda.synchronize()
    report["sort_ms"] = {"torch.sort of the int32 cell key": round(ev[0].elapsed_time(ev[1]), 3),
                         "gather of x, y, z, q through the permutation": round(ev[1].elapsed_time(ev[2]), 3)}

    def run(label, p):
        px, py, pz, pq = (a[p] for a in (x, y, z, q)) if p is not perm else (sx, sy, sz, sq)
        report[label] = stage_times(scb, mesh, px, py, pz, pq, outs)
        worst = max(float((o - r[p]).abs().max() / r.abs().max()) for o, r in zip(outs, ref))
        report[label]["max_rel_diff_vs_random_order"] = worst
        del px, py, pz, pq

    run("cell_sorted", perm)
    del sx, sy, sz, sq
    run("z_plane_sorted", torch.sort(iz.to(torch.int32)).indices)
    del iz, key
    for window in (1 << 12, 1 << 16, 1 << 20):
        # order degraded since the last sort: particles shuffled inside windows of `window` consecutive sorted particles
        n_full = (npart // window) * window
        noise = torch.rand(n_full, device="cuda").view(-1, window)
        local = torch.argsort(noise, dim=1) + (torch.arange(n_full // window, device="cuda") * window).view(-1, 1)
        p = torch.cat([perm[local.view(-1)], perm[n_full:]])
        del noise, local
        run("cell_sorted_then_shuffled_in_windows_of_%d" % window, p)
        del p
    print(json.dumps(report, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sorted_input_probe_%s.json" % dt), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
