"""Independent pin of the oracle's solve: E_c[p] = FPEI * sum_n rho[n] * IGF_c((p-n) delta + offset)
evaluated by direct summation with the IGF computed in 50-digit arithmetic (mpmath), so neither the
FFT structure nor double-precision cancellation in the 8-point differencing is shared with the
oracle.  Also checks the NumPy model of the GPU's restructured algorithm (tests/fused_model.py)."""
import mpmath as mp
import numpy as np
import pytest

from oracle import spacecharge_oracle as so

mp.mp.dps = 50


def F_mp(x, y, z):
    r = mp.sqrt(x * x + y * y + z * z)
    return x * mp.atan((y * z) / (r * x)) - z * mp.log(r + y) + y * mp.log((r - z) / (r + z)) / 2


def igf_mp(d, delta, gamma, icomp, off):
    dx, dy, dz = mp.mpf(float(delta[0])), mp.mpf(float(delta[1])), mp.mpf(float(delta[2])) * mp.mpf(float(gamma))
    fac = (mp.mpf(float(gamma)) if icomp in (1, 2) else mp.mpf(1)) / (dx * dy * dz)
    tot = mp.mpf(0)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                u = (d[0] - mp.mpf(1) / 2 + a) * dx + mp.mpf(float(off[0]))
                v = (d[1] - mp.mpf(1) / 2 + b) * dy + mp.mpf(float(off[1]))
                w = (d[2] - mp.mpf(1) / 2 + c) * dz + mp.mpf(float(off[2])) * mp.mpf(float(gamma))
                args = {1: (u, v, w), 2: (v, w, u), 3: (w, u, v)}[icomp]
                tot += (-1) ** (3 - a - b - c) * F_mp(*args)
    return tot * fac


@pytest.mark.parametrize("at_cathode", [False, True])
def test_fft_convolution_equals_direct_sum(at_cathode):
    grid = (5, 6, 4)
    rng = np.random.default_rng(17)
    m = so.mesh_from_bounds(grid, (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), gamma=3.0)
    m.rho[...] = rng.standard_normal(grid)
    so.solve(m, at_cathode=at_cathode)
    offz = 2 * m.min_bounds[2] + (m.max_bounds[2] - m.min_bounds[2])
    pts = [(0, 0, 0), (4, 5, 3), (2, 3, 1), (1, 4, 2)]
    for ic in (1, 2, 3):
        scale = np.abs(m.efield[..., ic - 1]).max()
        for p in pts:
            tot = mp.mpf(0)
            for n in np.ndindex(*grid):
                d = tuple(pi - ni for pi, ni in zip(p, n))
                tot += mp.mpf(float(m.rho[n])) * igf_mp(d, m.delta, m.gamma, ic, (0, 0, 0))
                if at_cathode:  # image of node n: charge -rho at mirrored z index, displaced by offz
                    nm = (n[0], n[1], grid[2] - 1 - n[2])
                    dm = tuple(pi - ni for pi, ni in zip(p, nm))
                    tot += -mp.mpf(float(m.rho[n])) * igf_mp(dm, m.delta, m.gamma, ic, (0, 0, offz))
            want = float(tot * mp.mpf(so.FPEI))
            assert abs(m.efield[p + (ic - 1,)] - want) < 2e-10 * scale, (ic, p)


@pytest.mark.parametrize("grid", [(6, 10, 5), (8, 8, 8), (3, 2, 4)])
@pytest.mark.parametrize("at_cathode", [False, True])
def test_restructured_algorithm_equals_reference_structure(grid, at_cathode):
    import fused_model as fm
    rng = np.random.default_rng(0)
    m = so.mesh_from_bounds(grid, (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), gamma=3.0)
    m.rho[...] = rng.standard_normal(grid)
    so.solve(m, at_cathode=at_cathode)
    e = fm.solve_fused(m.rho, m.delta, m.gamma, m.min_bounds[2], m.max_bounds[2], at_cathode)
    for c in range(3):
        assert np.abs(e[..., c] - m.efield[..., c]).max() / np.abs(m.efield[..., c]).max() < 1e-11
