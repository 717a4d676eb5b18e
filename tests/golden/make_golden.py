"""Regenerates the frozen fixtures in this directory from the oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot be run here (no Julia), so these are *oracle* outputs
(NumPy restatement, Float64) frozen so that a change in either the oracle or the CUDA path is noticed.
Small on purpose (< 100 kB in total)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import spacecharge_oracle as so  # noqa: E402


def main():
    # 1. single point charge, the fixtures quoted in SURVEY.md 8(c)
    out = {}
    for name, grid, gamma in (("g8_gamma1", (8, 8, 8), 1.0), ("g16_gamma2", (16, 16, 16), 2.0), ("g8_gamma10", (8, 8, 8), 10.0)):
        p = (np.array([0.0]),) * 3
        m, _ = so.full_step(grid, *p, np.array([1.0]), gamma=gamma)
        out[name + "_E211"] = m.efield[1, 0, 0, :].copy()
        out[name + "_E112"] = m.efield[0, 0, 1, :].copy()
        out[name + "_E222"] = m.efield[1, 1, 1, :].copy()
    np.savez(os.path.join(HERE, "single_charge.npz"), **out)

    # 2. random rho on an anisotropic non-power-of-two grid, free space and cathode
    grid, lo, hi, gamma = (6, 10, 5), (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), 3.0
    rho = np.random.default_rng(2024).standard_normal(grid)
    m = so.mesh_from_bounds(grid, lo, hi, gamma=gamma)
    m.rho[...] = rho
    so.solve(m)
    e_free = m.efield.copy()
    so.solve(m, at_cathode=True)
    np.savez(os.path.join(HERE, "random_rho_6x10x5.npz"), rho=rho, lo=lo, hi=hi, gamma=gamma, e_free=e_free,
             e_cathode=m.efield.copy())

    # 3. a small Gaussian bunch through the whole step
    rng = np.random.default_rng(7)
    n = 2000
    x, y, z = (rng.standard_normal(n) * s for s in (1e-3, 0.7e-3, 1.3e-3))
    q = rng.uniform(0.5, 1.5, n) * 1e-12
    grid = (12, 9, 15)
    m, (ex, ey, ez) = so.full_step(grid, x, y, z, q, gamma=1.5)
    np.savez(os.path.join(HERE, "gaussian_step_12x9x15.npz"), x=x, y=y, z=z, q=q, gamma=1.5, rho=m.rho, efield=m.efield,
             ex=ex, ey=ey, ez=ez, lo=m.min_bounds, hi=m.max_bounds, delta=m.delta)


if __name__ == "__main__":
    main()
