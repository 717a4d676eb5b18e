"""Seeded random sweep over grid shapes, spacings, gamma and boundary condition: the CUDA solve and the
whole step against the oracle.  Catches shape-dependent indexing mistakes (padding to the next power of
two, pitch handling, pruned stores, symmetric Green-spectrum build) that fixed test shapes can miss."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("seed", range(16))
def test_random_geometry_solve_and_step(scb, oracle, record, seed):
    import torch
    rng = np.random.default_rng(1000 + seed)
    grid = tuple(int(v) for v in rng.integers(2, 37, size=3))
    gamma = float(rng.choice([1.0, 1.7, 5.0, 30.0]))
    cath = bool(rng.integers(0, 2))
    T = np.float64 if seed % 4 else np.float32
    n = int(rng.integers(1, 4000))
    sig = rng.uniform(0.2e-3, 3e-3, size=3)
    x, y, z = (rng.standard_normal(n) * s for s in sig)
    z = z + (8e-3 if cath else 0.0)
    q = rng.uniform(-1.0, 2.0, n) * 1e-12
    ref, want = oracle.full_step(grid, x, y, z, q, T=np.float64, gamma=gamma, at_cathode=cath)
    d = [torch.from_numpy(a).cuda() for a in (x, y, z, q)]
    mesh = scb.Mesh3D(grid, *d[:3], T=T, gamma=gamma)
    if T == np.float32:   # grade Float32 against the Float64 oracle on the same Float32-valued geometry
        ref = oracle.mesh_from_particles(grid, x, y, z, T=np.float64, gamma=gamma)
        ref.min_bounds, ref.max_bounds, ref.delta = (tuple(np.float64(v) for v in t)
                                                     for t in (mesh.min_bounds, mesh.max_bounds, mesh.delta))
        oracle.deposit(ref, x, y, z, q, clamp=True)
        oracle.solve(ref, at_cathode=cath)
        want = oracle.interpolate_field(ref, x, y, z, clamp=True)
    scb.deposit_(mesh, *d)
    scb.solve_(mesh, at_cathode=cath)
    got = scb.interpolate_field(mesh, *d[:3])
    tol = 1e-10 if T == np.float64 else 2e-5
    e = mesh.efield.cpu().numpy()
    scale = max(np.abs(ref.efield[..., c]).max() for c in range(3))
    for c in range(3):
        # components that vanish by symmetry are compared on the scale of the field
        err = float(np.abs(e[..., c] - ref.efield[..., c]).max() / scale)
        record("E%d grid=%s gamma=%g cath=%s %s" % (c, grid, gamma, cath, np.dtype(T).name), err, tol)
        assert err < tol
    wscale = max(np.abs(w).max() for w in want)
    for c in range(3):
        assert float(np.abs(got[c].cpu().numpy() - want[c]).max() / wscale) < tol


@pytest.mark.parametrize("grid", [(1024, 2, 3), (2, 1024, 2), (3, 2, 1024), (512, 4, 2), (2, 2, 2), (129, 3, 65),
                                  (5, 3, 200), (4, 6, 256), (9, 2, 129)])
@pytest.mark.parametrize("at_cathode", [False, True])
def test_extreme_aspect_ratios(scb, oracle, record, grid, at_cathode):
    """Largest supported axis (n = 1024, padded transform length 2048), the smallest grid (2,2,2),
    sizes just above a power of two (129 -> padded 512), and 129..256 points along z (padded 512: the
    free-space solve takes the even/odd-bin z pass k_z_eo)."""
    import torch
    rng = np.random.default_rng(sum(grid))
    rho = rng.standard_normal(grid)
    lo, hi = (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3)
    ref = oracle.mesh_from_bounds(grid, lo, hi, gamma=2.0)
    ref.rho[...] = rho
    oracle.solve(ref, at_cathode=at_cathode)
    mesh = scb.Mesh3D(grid, lo, hi, gamma=2.0)
    mesh.rho.copy_(torch.from_numpy(rho).cuda())
    scb.solve_(mesh, at_cathode=at_cathode)
    e = mesh.efield.cpu().numpy()
    scale = max(np.abs(ref.efield[..., c]).max() for c in range(3))
    for c in range(3):
        err = float(np.abs(e[..., c] - ref.efield[..., c]).max() / scale)
        record("E%d grid=%s cath=%s" % (c, grid, at_cathode), err, 1e-10)
        assert err < 1e-10


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_even_odd_z_pass_line_by_line(scb, oracle, record, T, tol):
    """k_z_eo (padded z length 512): every kx column of a CTA tile, both ky halves of the folded spectrum, the
    last partial tile (kx = 256), nz < 256 (zero rows inside the 256-point transforms), with and without the
    potential as fourth component."""
    import torch
    grid = (130, 5, 200)   # Lx = 512: kx tiles up to the single-column tile at kx = 256
    rng = np.random.default_rng(5)
    rho = rng.standard_normal(grid).astype(T)
    lo, hi = (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3)
    mesh = scb.Mesh3D(grid, lo, hi, T=T, gamma=1.5)
    ref = oracle.mesh_from_bounds(grid, lo, hi, T=np.float64, gamma=1.5)
    if T == np.float32:   # the Float64 oracle on the same Float32-valued geometry
        ref.min_bounds, ref.max_bounds, ref.delta = (tuple(np.float64(v) for v in t)
                                                     for t in (mesh.min_bounds, mesh.max_bounds, mesh.delta))
    ref.rho[...] = rho
    oracle.solve(ref, potential=True)
    mesh.rho.copy_(torch.from_numpy(rho).cuda())
    scb.solve_(mesh)
    e = mesh.efield.cpu().numpy().astype(np.float64)
    scale = max(np.abs(ref.efield[..., c]).max() for c in range(3))
    for c in range(3):
        err = float(np.abs(e[..., c] - ref.efield[..., c]).max() / scale)
        record("k_z_eo E%d %s" % (c, np.dtype(T).name), err, tol)
        assert err < tol
    scb.solve_potential_(mesh)
    e2 = mesh.efield.cpu().numpy().astype(np.float64)
    for c in range(3):
        assert float(np.abs(e2[..., c] - ref.efield[..., c]).max() / scale) < tol
    perr = float(np.abs(mesh.phi.cpu().numpy() - ref.phi).max() / np.abs(ref.phi).max())
    # the potential's Green function cancels harder than the field's on this elongated geometry (parity unpinned
    # extension): same 1.7e-10 with the TMA z pass (SCB_Z_EO=0)
    record("k_z_eo phi %s" % np.dtype(T).name, perr, 5 * tol)
    assert perr < 5 * tol
