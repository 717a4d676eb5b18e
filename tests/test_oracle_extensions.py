"""Pins for the oracle's EXTENSION functions (potential, magnetic field, fused kick -- SURVEY.md 8(f)-2/3).

The reference has no phi / B / kick output (src/mesh.jl:19-34 holds rho and efield only), so these are
"parity unpinned" against the reference by construction; they are pinned here against mathematics instead:
direct summation with the potential IGF in 50-digit arithmetic, the analytic potential of an isotropic
Gaussian, and the relation E = -grad(phi) that ties the extension to the reference's own field components.
"""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import spacecharge_oracle as so

mp.mp.dps = 50


def P_mp(x, y, z):  # src/green_functions.jl:13-22
    r = mp.sqrt(x * x + y * y + z * z)
    return (-z * z * mp.atan(x * y / (z * r)) / 2 - y * y * mp.atan(x * z / (y * r)) / 2 - x * x * mp.atan(y * z / (x * r)) / 2
            + y * z * mp.log(x + r) + x * z * mp.log(y + r) + x * y * mp.log(z + r))


def igf_phi_mp(d, delta, gamma, off):
    dx, dy, dz = mp.mpf(float(delta[0])), mp.mpf(float(delta[1])), mp.mpf(float(delta[2])) * mp.mpf(float(gamma))
    tot = mp.mpf(0)
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                u = (d[0] - mp.mpf(1) / 2 + a) * dx + mp.mpf(float(off[0]))
                v = (d[1] - mp.mpf(1) / 2 + b) * dy + mp.mpf(float(off[1]))
                w = (d[2] - mp.mpf(1) / 2 + c) * dz + mp.mpf(float(off[2])) * mp.mpf(float(gamma))
                tot += (-1) ** (3 - a - b - c) * P_mp(u, v, w)
    return tot / (dx * dy * dz)


@pytest.mark.parametrize("at_cathode", [False, True])
def test_potential_equals_direct_sum(at_cathode):
    grid = (5, 6, 4)
    rng = np.random.default_rng(23)
    m = so.mesh_from_bounds(grid, (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), gamma=3.0)
    m.rho[...] = rng.standard_normal(grid)
    so.solve(m, at_cathode=at_cathode, potential=True)
    offz = 2 * m.min_bounds[2] + (m.max_bounds[2] - m.min_bounds[2])
    scale = np.abs(m.phi).max()
    for p in [(0, 0, 0), (4, 5, 3), (2, 3, 1), (1, 4, 2)]:
        tot = mp.mpf(0)
        for n in np.ndindex(*grid):
            d = tuple(pi - ni for pi, ni in zip(p, n))
            tot += mp.mpf(float(m.rho[n])) * igf_phi_mp(d, m.delta, m.gamma, (0, 0, 0))
            if at_cathode:
                nm = (n[0], n[1], grid[2] - 1 - n[2])
                dm = tuple(pi - ni for pi, ni in zip(p, nm))
                tot += -mp.mpf(float(m.rho[n])) * igf_phi_mp(dm, m.delta, m.gamma, (0, 0, offz))
        want = float(tot * mp.mpf(so.FPEI))
        assert abs(m.phi[p] - want) < 2e-10 * scale, p


def test_potential_of_isotropic_gaussian():
    """phi(r) = Q/(4 pi eps0) erf(r / (sqrt2 sigma)) / r, same bunch as test/analytical_test.jl:20-30."""
    rng = np.random.default_rng(123)
    n, sigma, Q = 1_000_000, 1e-3, 1e-9
    x, y, z = (rng.standard_normal(n) * sigma for _ in range(3))
    q = np.full(n, Q / n)
    m = so.mesh_from_particles((32, 32, 32), x, y, z)
    so.deposit(m, x, y, z, q)
    so.solve(m, potential=True)
    ax = [m.min_bounds[a] + m.delta[a] * np.arange(32) for a in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    erf = np.vectorize(math.erf)
    want = so.FPEI * Q * erf(r / (math.sqrt(2) * sigma)) / r
    err = np.abs(m.phi - want).max() / np.abs(want).max()
    assert err < 0.02, err


@pytest.mark.parametrize("gamma", [1.0, 4.0])
def test_field_is_minus_gradient_of_potential(gamma):
    """Ties the extension to the reference's own components: on the lab-frame mesh
    E_{x,y} = -gamma d(phi)/d{x,y} and E_z = -(1/gamma) d(phi)/dz (rest-frame potential, dz' = gamma dz)."""
    rng = np.random.default_rng(5)
    n, sigma = 400_000, 1e-3
    x, y = (rng.standard_normal(n) * sigma for _ in range(2))
    z = rng.standard_normal(n) * sigma / gamma          # a bunch that is round in its rest frame
    q = np.full(n, 1e-9 / n)
    m = so.mesh_from_particles((24, 24, 24), x, y, z, gamma=gamma)
    so.deposit(m, x, y, z, q)
    so.solve(m, potential=True)
    sl = slice(1, -1)
    ex = -gamma * (m.phi[2:, sl, sl] - m.phi[:-2, sl, sl]) / (2 * m.delta[0])
    ey = -gamma * (m.phi[sl, 2:, sl] - m.phi[sl, :-2, sl]) / (2 * m.delta[1])
    ez = -(m.phi[sl, sl, 2:] - m.phi[sl, sl, :-2]) / (2 * m.delta[2]) / gamma
    for c, fd in enumerate((ex, ey, ez)):
        e = m.efield[sl, sl, sl, c]
        assert np.abs(fd - e).max() / np.abs(e).max() < 0.05, c


def test_magnetic_field_and_kick():
    rng = np.random.default_rng(2)
    m = so.mesh_from_bounds((6, 5, 7), (-1, -1, -1), (1, 1, 1), gamma=5.0)
    m.efield[...] = rng.standard_normal(m.efield.shape)
    b = so.magnetic_field(m)
    beta = math.sqrt(1 - 1 / 25.0)
    assert np.allclose(b[..., 0], -beta / so.CLIGHT * m.efield[..., 1], rtol=1e-15)
    assert np.allclose(b[..., 1], beta / so.CLIGHT * m.efield[..., 0], rtol=1e-15)
    assert not b[..., 2].any()
    x, y, z = (rng.uniform(-0.9, 0.9, 100) for _ in range(3))
    p0 = [rng.standard_normal(100) for _ in range(3)]
    e = so.interpolate_field(m, x, y, z)
    got = so.interpolate_kick(m, x, y, z, *p0, 0.25, -2.0)
    for g, p, ec, c in zip(got, p0, e, (0.25, 0.25, -2.0)):
        assert np.array_equal(g, p + c * ec)
