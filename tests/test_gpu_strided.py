"""Strided / array-of-structures particle layouts (scb_*_strided): the records are read and
written in place, and every result must be bit-identical to the dense (reference-layout) call on the same
values -- the kernels are the same template with the element stride compiled in or out.  The dense calls are
the ones graded against the oracle in test_gpu_parity.py; one oracle check here closes the loop directly."""
import numpy as np
import pytest

from test_gpu_parity import TOL64, check, gaussian, to_dev

pytestmark = pytest.mark.gpu


def records(x, y, z, extra=3, seed=5):
    """(Np, 3 + extra) phase-space-like records: columns 0, 2, 4 = x, y, z; the others momenta / padding."""
    import torch
    rng = np.random.default_rng(seed)
    rec = rng.standard_normal((len(x), 3 + extra)).astype(x.dtype)
    rec[:, 0], rec[:, 2], rec[:, 4] = x, y, z
    return torch.from_numpy(rec).cuda()


@pytest.mark.parametrize("pdt,mdt", [(np.float64, np.float64), (np.float32, np.float32),
                                     (np.float64, np.float32), (np.float32, np.float64)])
@pytest.mark.parametrize("npart,grid", [(300_000, (24, 20, 28)), (3_000, (24, 20, 28))])   # tile deposit + packed gather / small-bunch kernels
def test_aos_views_match_dense_bitwise(scb, pdt, mdt, npart, grid):
    import torch
    x, y, z, q = gaussian(npart, 21, dtype=pdt)
    dx, dy, dz, dq = to_dev(x, y, z, q)
    rec = records(x, y, z)
    vx, vy, vz = rec[:, 0], rec[:, 2], rec[:, 4]
    assert vx.stride(0) == 6 and not vx.is_contiguous()

    dense = scb.Mesh3D(grid, dx, dy, dz, T=mdt)
    aos = scb.Mesh3D(grid, vx, vy, vz, T=mdt)                  # scb_bounds_strided
    assert (aos.min_bounds, aos.max_bounds, aos.delta) == (dense.min_bounds, dense.max_bounds, dense.delta)

    scb.deposit_(dense, dx, dy, dz, dq)
    scb.deposit_(aos, vx, vy, vz, dq)                          # scb_deposit_strided
    # reductions into rho: the accumulation order is not fixed, so rho agrees to round-off, not bitwise
    tol = 1e-12 if mdt == np.float64 else 1e-5
    assert float((aos.rho - dense.rho).abs().max() / dense.rho.abs().max()) < tol
    aos.rho.copy_(dense.rho)

    scb.solve_(dense)
    scb.solve_(aos)
    assert bool((aos.efield == dense.efield).all())

    want = scb.interpolate_field(dense, dx, dy, dz)
    got = scb.interpolate_field(aos, vx, vy, vz)               # scb_interpolate_strided
    for a, b in zip(got, want):
        assert a.is_contiguous() and torch.equal(a, b)
    # the records were only read
    assert torch.equal(rec[:, 0], dx) and torch.equal(rec[:, 2], dy) and torch.equal(rec[:, 4], dz)


def test_broadcast_charge_and_strided_charge(scb):
    """stride_q = 0: one charge for every particle (equal-weight macro-particles); stride_q > 1: charge inside
    the record."""
    import torch
    n, grid = 200_000, (16, 16, 16)
    x, y, z, q = gaussian(n, 3)
    dx, dy, dz, dq = to_dev(x, y, z, q)
    dense = scb.Mesh3D(grid, dx, dy, dz)
    scb.deposit_(dense, dx, dy, dz, dq)
    m2 = scb.Mesh3D(grid, dx, dy, dz)
    q0 = dq[:1].expand(n)
    assert q0.stride(0) == 0
    scb.deposit_(m2, dx, dy, dz, q0)
    assert float((m2.rho - dense.rho).abs().max() / dense.rho.abs().max()) < 1e-12
    assert abs(float(m2.rho.sum()) - q.sum()) < 1e-12 * abs(q.sum())
    rec = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = dx, dy, dz, dq
    m3 = scb.Mesh3D(grid, rec[:, 0], rec[:, 1], rec[:, 2])
    scb.deposit_(m3, rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3])
    assert float((m3.rho - dense.rho).abs().max() / dense.rho.abs().max()) < 1e-12


@pytest.mark.parametrize("pdt,mdt", [(np.float64, np.float64), (np.float32, np.float32)])
def test_kick_in_place_on_phase_space_records(scb, pdt, mdt):
    """Bmad-style records (x, px, y, py, z, pz): the fused kick updates px, py, pz inside the records; the
    coordinates stay untouched and the momenta equal the dense kick bit for bit."""
    import torch
    n, grid = 250_000, (20, 24, 18)
    x, y, z, q = gaussian(n, 77, dtype=pdt)
    dx, dy, dz, dq = to_dev(x, y, z, q)
    rec = records(x, y, z)
    before = rec.clone()
    mesh = scb.Mesh3D(grid, dx, dy, dz, T=mdt)
    scb.deposit_(mesh, dx, dy, dz, dq)
    scb.solve_(mesh)
    px, py, pz = (before[:, c].clone() for c in (1, 3, 5))
    scb.interpolate_kick_(mesh, dx, dy, dz, px, py, pz, 1.5e-9, -2.5e-9)
    scb.interpolate_kick_(mesh, rec[:, 0], rec[:, 2], rec[:, 4], rec[:, 1], rec[:, 3], rec[:, 5], 1.5e-9, -2.5e-9)
    assert torch.equal(rec[:, 1], px) and torch.equal(rec[:, 3], py) and torch.equal(rec[:, 5], pz)
    for c in (0, 2, 4):
        assert torch.equal(rec[:, c], before[:, c])
    assert not torch.equal(rec[:, 1], before[:, 1])


def test_step_on_records_matches_oracle(scb, oracle, record):
    """Whole step on AoS records with strided outputs (E written into the records), against the oracle."""
    import torch
    n, grid = 120_000, (16, 20, 12)
    x, y, z, q = gaussian(n, 9)
    ref = oracle.mesh_from_particles(grid, x, y, z)
    oracle.deposit(ref, x, y, z, q, clamp=True)
    oracle.solve(ref)
    want = oracle.interpolate_field(ref, x, y, z, clamp=True)
    rec = torch.zeros((n, 7), dtype=torch.float64, device="cuda")     # x, y, z, q, Ex, Ey, Ez
    for c, a in enumerate((x, y, z, q)):
        rec[:, c] = torch.from_numpy(a).cuda()
    mesh = scb.Mesh3D(grid, rec[:, 0], rec[:, 1], rec[:, 2])
    assert mesh.delta == ref.delta and mesh.min_bounds == ref.min_bounds
    scb.step_(mesh, rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5], rec[:, 6])
    check(record, "rho (records)", mesh.rho.cpu().numpy(), ref.rho, TOL64)
    for c in range(3):
        check(record, "E%d at particles (records)" % c, rec[:, 4 + c].cpu().numpy(), want[c], TOL64)
    assert torch.equal(rec[:, 0].cpu(), torch.from_numpy(x))


def test_strided_argument_errors(scb):
    import ctypes as C
    import torch
    mesh = scb.Mesh3D((8, 8, 8), (-1, -1, -1), (1, 1, 1))
    hd, lib = mesh.handle, mesh.handle.lib
    x = torch.zeros(16, dtype=torch.float64, device="cuda")
    e = torch.zeros(16, dtype=torch.float64, device="cuda")
    for bad in (scb._strides(0, 1, 1), scb._strides(1, -2, 1), scb._strides(1, 1, 1, -1)):
        rc = lib.scb_deposit_strided(hd.h, 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), C.byref(bad), 1,
                                     mesh._rho.data_ptr(), 1, mesh._n(), mesh._lo(), mesh._d(), 1)
        assert rc == -1
    rc = lib.scb_interpolate_strided(hd.h, 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), C.byref(scb._strides(1, 1, 1, 1, 0, 1, 1)),
                                     1, mesh._efield.data_ptr(), 1, mesh._n(), mesh._lo(), mesh._d(), e.data_ptr(),
                                     e.data_ptr(), e.data_ptr())
    assert rc == -1
    rc = lib.scb_deposit_strided(hd.h, 4, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), None, 1,
                                 mesh._rho.data_ptr(), 1, mesh._n(), mesh._lo(), mesh._d(), 1)
    assert rc == -1
    assert "null strides" in lib.scb_last_error(hd.h).decode()
