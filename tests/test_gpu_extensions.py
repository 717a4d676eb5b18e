"""GPU parity of the extension entry points with the oracle's restatement:
scb_solve_potential (phi through the fused passes as a fourth component), scb_bfield, scb_interpolate_kick.
The reference has no counterpart (src/mesh.jl:19-34 stores rho and efield only), so the oracle functions
used here are themselves pinned by tests/test_oracle_extensions.py (direct summation, analytic Gaussian,
E = -grad phi)."""
import math

import numpy as np
import pytest

from test_gpu_parity import TOL32, TOL64, check, gaussian, set_rho, to_dev

pytestmark = pytest.mark.gpu


def _pair(scb, oracle, grid, lo, hi, gamma, rho, T=np.float64):
    ref = oracle.mesh_from_bounds(grid, lo, hi, T=np.float64, gamma=gamma)
    mesh = scb.Mesh3D(grid, lo, hi, T=T, gamma=gamma)
    if T == np.float32:
        ref.min_bounds, ref.max_bounds, ref.delta = (tuple(np.float64(v) for v in t)
                                                     for t in (mesh.min_bounds, mesh.max_bounds, mesh.delta))
    ref.rho[...] = rho.astype(T)
    set_rho(mesh, rho.astype(T))
    return mesh, ref


@pytest.mark.parametrize("grid", [(8, 8, 8), (6, 10, 5), (2, 3, 4), (33, 17, 40), (32, 32, 32)])
@pytest.mark.parametrize("at_cathode", [False, True])
def test_potential_matches_oracle_f64(scb, oracle, record, grid, at_cathode):
    rng = np.random.default_rng(sum(grid) + 1)
    rho = rng.standard_normal(grid)
    mesh, ref = _pair(scb, oracle, grid, (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), 3.0, rho)
    oracle.solve(ref, at_cathode=at_cathode, potential=True)
    scb.solve_potential_(mesh, at_cathode=at_cathode)
    check(record, "phi", mesh.phi.cpu().numpy(), ref.phi, TOL64)
    got = mesh.efield.cpu().numpy()
    for c in range(3):
        check(record, "E%d" % c, got[..., c], ref.efield[..., c], TOL64)


def test_potential_f32_and_cache_upgrade(scb, oracle, record):
    """A plain solve_ first (3-component spectrum cached), then solve_potential_ on the same geometry:
    the cached entry must be rebuilt with the potential component, and E must not change."""
    grid = (16, 24, 32)
    rng = np.random.default_rng(8)
    rho = rng.standard_normal(grid)
    mesh, ref = _pair(scb, oracle, grid, (-1e-3, -2e-3, 0.5e-3), (1e-3, 1.5e-3, 2.5e-3), 2.0, rho, T=np.float32)
    oracle.solve(ref, potential=True)
    scb.solve_(mesh)
    e0 = mesh.efield.clone()
    scb.solve_potential_(mesh)
    assert bool((mesh.efield == e0).all())
    check(record, "phi f32", mesh.phi.cpu().numpy(), ref.phi, TOL32)
    scb.solve_(mesh)       # and back: the 4-component entry serves the plain solve
    assert bool((mesh.efield == e0).all())
    with pytest.raises(scb.ErrorException):
        scb.Mesh3D(grid, (-1, -1, -1), (1, 1, 1)).phi


def test_potential_of_gaussian_bunch_on_gpu(scb, record):
    """Full path (deposit + solve_potential_) against the analytic potential of an isotropic Gaussian."""
    n, sigma, Q = 1_000_000, 1e-3, 1e-9
    x, y, z, q = gaussian(n, 123, sigma, Q)
    d = to_dev(x, y, z, q)
    mesh = scb.Mesh3D((64, 64, 64), *d[:3])
    scb.deposit_(mesh, *d)
    scb.solve_potential_(mesh)
    ax = [mesh.min_bounds[a] + mesh.delta[a] * np.arange(64) for a in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    want = scb.FPEI * Q * np.vectorize(math.erf)(r / (math.sqrt(2) * sigma)) / r
    err = np.abs(mesh.phi.cpu().numpy() - want).max() / np.abs(want).max()
    record("phi vs analytic Gaussian 64^3", err, 0.02)
    assert err < 0.02


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_magnetic_field(scb, oracle, T):
    import torch
    grid = (9, 7, 11)
    rng = np.random.default_rng(4)
    e = rng.standard_normal(grid + (3,)).astype(T)
    ref = oracle.mesh_from_bounds(grid, (-1, -1, -1), (1, 1, 1), T=T, gamma=7.0)
    mesh = scb.Mesh3D(grid, (-1, -1, -1), (1, 1, 1), T=T, gamma=7.0)
    ref.efield[...] = e
    mesh.efield.copy_(torch.from_numpy(e).cuda())
    got = scb.magnetic_field(mesh).cpu().numpy()
    assert np.array_equal(got, oracle.magnetic_field(ref))


@pytest.mark.parametrize("pdt,mdt,tol", [(np.float64, np.float64, 1e-13), (np.float32, np.float32, TOL32),
                                         (np.float64, np.float32, TOL32), (np.float32, np.float64, 1e-6)])
@pytest.mark.parametrize("n", [50000, 300])          # packed-field gather / 24-gather kernel
def test_interpolate_kick_matches_oracle(scb, oracle, record, pdt, mdt, tol, n):
    import torch
    x, y, z, q = gaussian(n, 19, dtype=pdt)
    grid = (20, 31, 16)
    ref = oracle.mesh_from_particles(grid, x, y, z, T=mdt)
    mesh = scb.Mesh3D(grid, *to_dev(x, y, z), T=mdt)
    rng = np.random.default_rng(6)
    e = rng.standard_normal(grid + (3,)).astype(mdt)
    ref.efield[...] = e
    mesh.efield.copy_(torch.from_numpy(e).cuda())
    p0 = [rng.standard_normal(n).astype(pdt) for _ in range(3)]
    want = oracle.interpolate_kick(ref, x, y, z, *p0, 0.37, -1.9, clamp=True)
    mom = list(to_dev(*p0))
    scb.interpolate_kick_(mesh, *to_dev(x, y, z), *mom, 0.37, -1.9)
    for g, w, p in zip(mom, want, p0):
        assert g.dtype == (torch.float32 if pdt == np.float32 else torch.float64)
        # graded on the increment, which is what the kernel computes
        check(record, "kick", g.cpu().numpy().astype(np.float64) - p, w.astype(np.float64) - p, max(tol, 1e-6 if pdt == np.float32 else tol))
    # the kick with the same coefficients equals interpolate_field followed by the update
    ex, ey, ez = scb.interpolate_field(mesh, *to_dev(x, y, z))
    mom2 = list(to_dev(*p0))
    scb.interpolate_kick_(mesh, *to_dev(x, y, z), *mom2, 1.0, 1.0)
    for g, p, ec in zip(mom2, p0, (ex, ey, ez)):
        assert np.array_equal(g.cpu().numpy(), (p.astype(np.float64) + ec.cpu().numpy().astype(np.float64)).astype(pdt)) or \
            np.abs(g.cpu().numpy() - (p + ec.cpu().numpy())).max() <= 4 * np.finfo(pdt).eps * np.abs(p + ec.cpu().numpy()).max()


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("at_cathode", [False, True])
def test_remesh_step_with_spectrum_prefetch_equals_separate_calls(scb, T, at_cathode):
    """Tracking loop: remesh_ + the fused step on a new geometry builds the Green spectrum on a
    second stream while the deposit runs (prefetch_green); the field must equal, bit for bit, what deposit_ / solve_ /
    interpolate_field give on a mesh constructed from scratch for the same particles and the same rho."""
    import torch
    n, grid = 150_000, (20, 16, 24)
    x, y, z, q = gaussian(n, 31, shift=(0, 0, 7e-3 if at_cathode else 0))
    dx, dy, dz, dq = to_dev(x, y, z, q)
    mesh = scb.Mesh3D(grid, dx, dy, dz, T=T, gamma=2.5)
    outs = [torch.empty_like(dx) for _ in range(3)]
    scb.step_(mesh, dx, dy, dz, dq, *outs, at_cathode=at_cathode)          # warm: geometry 0 cached
    for scale in (1.01, 0.97, 1.01):                                       # new, new, and back to a retired geometry
        xs = dx * scale
        mesh.remesh_(xs, dy, dz)
        scb.step_(mesh, xs, dy, dz, dq, *outs, at_cathode=at_cathode)
        fresh = scb.Mesh3D(grid, xs, dy, dz, T=T, gamma=2.5)
        assert (fresh.min_bounds, fresh.max_bounds, fresh.delta) == (mesh.min_bounds, mesh.max_bounds, mesh.delta)
        fresh.rho.copy_(mesh.rho)                                          # (reductions: rho itself is not bit-reproducible)
        fresh.handle.drop_green_cache()                                    # rebuilt inside solve_, on the main stream
        scb.solve_(fresh, at_cathode=at_cathode)
        want = scb.interpolate_field(fresh, xs, dy, dz)
        assert torch.equal(mesh.efield, fresh.efield)
        for a, b in zip(outs, want):
            assert torch.equal(a, b)


@pytest.mark.parametrize("at_cathode", [False, True])
def test_remesh_step_matches_the_oracle(scb, oracle, record, at_cathode):
    """remesh_ against the ORACLE (not against another CUDA mesh): after every re-fit the geometry must be what the
    reference's particle-based constructor computes for the new positions (src/mesh.jl:118-156, restated in
    oracle.mesh_from_particles) and the fused step on it must give the oracle's rho, E and interpolated E (1e-10)."""
    import torch
    n, grid = 60_000, (14, 18, 12)
    x, y, z, q = gaussian(n, 77, shift=(0, 0, 7e-3 if at_cathode else 0))
    dx, dy, dz, dq = to_dev(x, y, z, q)
    mesh = scb.Mesh3D(grid, dx, dy, dz, gamma=1.7)
    outs = [torch.empty_like(dx) for _ in range(3)]
    scb.step_(mesh, dx, dy, dz, dq, *outs, at_cathode=at_cathode)
    for k, scale in enumerate((1.03, 0.9, 1.2)):
        xs, ys = x * scale, y / scale
        dxs, dys = to_dev(xs, ys)
        mesh.remesh_(dxs, dys, dz)
        scb.step_(mesh, dxs, dys, dz, dq, *outs, at_cathode=at_cathode)
        ref, want = oracle.full_step(grid, xs, ys, z, q, gamma=1.7, at_cathode=at_cathode)
        assert (mesh.min_bounds, mesh.max_bounds, mesh.delta) == (ref.min_bounds, ref.max_bounds, ref.delta)
        check(record, "remesh %d rho" % k, mesh.rho.cpu().numpy(), ref.rho, TOL64)
        for c in range(3):
            check(record, "remesh %d E%d" % (k, c), mesh.efield[..., c].cpu().numpy(), ref.efield[..., c], TOL64)
            check(record, "remesh %d Einterp%d" % (k, c), outs[c].cpu().numpy(), want[c], TOL64)


def test_warm_step_is_capturable_in_a_cuda_graph(scb, record):
    """The small configurations are launch-bound (nine launches of 5-10 us at 32^3), so a tracking loop wants the
    step inside a CUDA graph.  In the warm state (workspace, packed field, Green spectrum built) scb_step performs no
    allocation, synchronisation or cross-stream wait, i.e. it is legal under stream capture: capture it once, replay it
    on new particle values in the same buffers, compare with direct calls."""
    import time
    import torch
    grid = (32, 32, 32)
    x, y, z, q = to_dev(*gaussian(100000, 21))
    lo = tuple(1.5 * float(a.min()) for a in (x, y, z))
    hi = tuple(1.5 * float(a.max()) for a in (x, y, z))
    mesh = scb.Mesh3D(grid, lo, hi, gamma=1.5)
    outs = [torch.empty_like(x) for _ in range(3)]

    def direct():
        scb.step_(mesh, x, y, z, q, *outs)
        torch.cuda.synchronize()
        return [o.clone() for o in outs], mesh.efield.clone()

    direct()
    want, e_want = direct()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        scb.step_(mesh, x, y, z, q, *outs)          # binds the handle to the capture stream, warm there
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            scb.step_(mesh, x, y, z, q, *outs)
    for o in outs:
        o.fill_(float("nan"))
    g.replay()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())   # noqa: E731
    for a, b in zip(outs, want):
        assert rel(a, b) < 1e-12      # atomics: rounding-level differences between runs
    assert rel(mesh.efield, e_want) < 1e-12
    # new particle values in the captured buffers (still inside the fixed mesh)
    x.mul_(0.7); y.mul_(-0.9); z.mul_(0.8)
    g.replay()
    torch.cuda.synchronize()
    got = [o.clone() for o in outs]
    want2, _ = direct()
    for a, b in zip(got, want2):
        assert rel(a, b) < 1e-12

    def per_step(fn, n=200):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n
    record("basic config, direct step (ms)", per_step(lambda: scb.step_(mesh, x, y, z, q, *outs)))
    record("basic config, graph replay (ms)", per_step(g.replay))
