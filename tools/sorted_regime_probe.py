"""Cell-ordered regime at BASELINE config 5 (1e8 particles, 256^3): what the sort costs, what the run-accumulating
kernels gain, how the gain decays while the bunch drifts, and the break-even number of steps per re-sort.

Orders timed (deposit / solve / interpolate minima over `reps` steps, CUDA events of the handle):
  random            the benchmark's order, default kernels                       (today's headline)
  random+cell       the same random bunch through the SCB_ORDER_CELL kernels     (what a wrong hint costs)
  sorted            scb_sort_particles + scb_permute, SCB_ORDER_CELL kernels
  sorted+default    the ordered bunch through the default kernels
  drift f           ordered bunch after every particle moved by N(0, (f*delta)^2) per axis without re-sorting
                    (f = fraction of a cell; a tracking step moves particles by much less than a cell)

usage: python tools/sorted_regime_probe.py [f64|f32] [particles] [grid]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

SIGMA, QTOT = 1e-3, 1e-9


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


def timed(fn, reps=3):
    best = None
    out = None
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


def main():
    dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
    npart = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
    g = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    grid = (g, g, g)
    scb = load_package()
    tdt = torch.float64 if dt == "f64" else torch.float32
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    x, y, z = (torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * SIGMA for _ in range(3))
    q = torch.full((npart,), QTOT / npart, device="cuda", dtype=tdt)
    mesh = scb.Mesh3D(grid, x, y, z, T=np.float64 if dt == "f64" else np.float32)
    outs = [torch.empty_like(x) for _ in range(3)]
    rep = {"dtype": dt, "particles": npart, "grid": list(grid)}
    rep["order_fraction_random"] = scb.particle_order_fraction(mesh, x, y, z)
    rep["random"] = stage_times(scb, mesh, x, y, z, q, outs)
    ref = [o.clone() for o in outs]
    rho_ref = mesh.rho.clone()
    scb.set_particle_order(mesh, "cell")
    rep["random+cell"] = stage_times(scb, mesh, x, y, z, q, outs, reps=2)
    scb.set_particle_order(mesh, "random")

    t_sort, perm = timed(lambda: scb.sort_particles(mesh, x, y, z))
    t_perm, srt = timed(lambda: scb.permute(perm, x, y, z, q, handle=mesh.handle))
    rep["sort_ms"] = {"scb_sort_particles": t_sort, "scb_permute(x,y,z,q)": t_perm, "total": round(t_sort + t_perm, 4)}
    sx, sy, sz, sq = srt
    rep["order_fraction_sorted"] = scb.particle_order_fraction(mesh, sx, sy, sz)
    rep["sorted+default"] = stage_times(scb, mesh, sx, sy, sz, sq, outs)
    scb.set_particle_order(mesh, "cell")
    rep["sorted"] = stage_times(scb, mesh, sx, sy, sz, sq, outs)
    pl = perm.long()
    rep["sorted"]["max_rel_diff_E_vs_random_order"] = max(
        float((o - r[pl]).abs().max() / r.abs().max()) for o, r in zip(outs, ref))
    rep["sorted"]["max_rel_diff_rho_vs_random_order"] = float((mesh.rho - rho_ref).abs().max() / rho_ref.abs().max())
    del pl, ref
    # re-sorting an already ordered bunch (the steady state of a tracking loop)
    t_resort, _ = timed(lambda: scb.sort_particles(mesh, sx, sy, sz))
    rep["sort_ms"]["scb_sort_particles on the ordered bunch"] = t_resort
    d = [float(v) for v in mesh.delta]
    for frac in (0.02, 0.1, 0.3, 1.0):
        gen.manual_seed(7)
        mx, my, mz = (s + torch.randn(npart, generator=gen, device="cuda", dtype=tdt) * (frac * dd)
                      for s, dd in zip((sx, sy, sz), d))
        # keep the mesh geometry: clamp the few particles that left the bounds
        for a, lo, hi in zip((mx, my, mz), mesh.min_bounds, mesh.max_bounds):
            a.clamp_(float(lo), float(hi))
        key = "drift %.2f cell" % frac
        rep[key] = stage_times(scb, mesh, mx, my, mz, sq, outs, reps=3)
        rep[key]["order_fraction"] = scb.particle_order_fraction(mesh, mx, my, mz)
        del mx, my, mz
    scb.set_particle_order(mesh, "random")
    gain = rep["random"]["step"] - rep["sorted"]["step"]
    rep["break_even_steps_per_sort"] = round(rep["sort_ms"]["total"] / gain, 2) if gain > 0 else None
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sorted_regime_probe_%s.json" % dt), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
