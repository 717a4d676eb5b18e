"""Cell-ordered bunches (scb_sort_particles / scb_permute / SCB_ORDER_CELL kernels) against the oracle.

The reference takes particles in any order (src/deposition.jl:218-247, src/interpolation.jl:100-128), so the results
of deposit! and interpolate_field do not depend on the order beyond the rounding of the (unordered) accumulation.  The
run-accumulating kernels must therefore (i) match the oracle for ANY order -- random, ordered, partly degraded -- and
(ii) agree with the default kernels; the sort itself is bit-exact integer work and is checked against a stable sort of
the oracle's cell keys.

Tolerances: rho and E max|a-b|/max|b| <= 1e-10 (Float64 meshes), <= 1e-5 (Float32 meshes, against the Float64 oracle
on the same Float32-valued geometry); interpolated values of k_interpolate_runs bit-identical to the oracle's
left-to-right sum in Float64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def gaussian(n, seed, sigma=1e-3, Q=1e-9, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x, y, z = ((rng.standard_normal(n) * sigma).astype(dtype) for _ in range(3))
    q = (rng.random(n) * (2 * Q / n)).astype(dtype)
    return x, y, z, q


def to_dev(*arrs):
    import torch
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs)


def oracle_keys(oracle, mesh, x, y, z):
    """linear cell index with the clamp of the CUDA kernels, from the oracle's index arithmetic"""
    W = np.float32 if (x.dtype == np.float32 and mesh.T == np.float32) else np.float64
    idx = []
    for p, lo, d, n in zip((x, y, z), mesh.min_bounds, mesh.delta, mesh.grid_size):
        i, _ = oracle.cell_index_and_frac(p.astype(W), W(lo), W(d))
        idx.append(np.clip(i, 0, n - 2))
    nx, ny, _ = mesh.grid_size
    return idx[0] + nx * (idx[1] + ny * idx[2])


@pytest.mark.parametrize("n,grid", [(1, (4, 4, 4)), (31, (4, 5, 6)), (8192, (16, 16, 16)), (8193, (16, 16, 16)),
                                    (100003, (37, 64, 19)), (3000000, (128, 128, 128)), (500000, (256, 256, 256)),
                                    (70000, (300, 7, 33))])
@pytest.mark.parametrize("pdt,mdt", [(np.float64, np.float64), (np.float32, np.float32)])
def test_sort_is_the_stable_order_by_cell_key(scb, oracle, n, grid, pdt, mdt):
    import torch
    x, y, z, q = gaussian(n, 5 + n % 97, dtype=pdt)
    dx, dy, dz = to_dev(x, y, z)
    mesh = scb.Mesh3D(grid, dx, dy, dz, T=mdt)
    perm = scb.sort_particles(mesh, dx, dy, dz).cpu().numpy().astype(np.int64)
    keys = oracle_keys(oracle, mesh, x, y, z)
    want = np.argsort(keys, kind="stable")
    assert np.array_equal(perm, want)
    # scb_permute applies it
    sx, sy, sz = scb.permute(torch.from_numpy(perm.astype(np.int32)).cuda(), dx, dy, dz, handle=mesh.handle)
    assert np.array_equal(sx.cpu().numpy(), x[want]) and np.array_equal(sz.cpu().numpy(), z[want])


def test_permute_many_fields_and_sort_particles_(scb):
    import torch
    x, y, z, q = gaussian(50000, 3)
    extra = [np.random.default_rng(k).random(50000) for k in range(7)]   # 11 arrays in all: two passes
    dev = to_dev(x, y, z, q, *extra)
    mesh = scb.Mesh3D((24, 24, 24), *dev[:3])
    out = scb.sort_particles_(mesh, *dev)
    perm = out[0].cpu().numpy().astype(np.int64)
    assert sorted(perm.tolist()) == list(range(50000))
    for got, src in zip(out[1:], (x, y, z, q, *extra)):
        assert np.array_equal(got.cpu().numpy(), src[perm])
    assert scb.particle_order_fraction(mesh, *out[1:4]) > 0.5
    assert scb.particle_order_fraction(mesh, *dev[:3]) < 0.2
    with pytest.raises(scb.ErrorException):
        scb.permute(out[0], dev[0], dev[1][:10])
    with pytest.raises(scb.ScbError):   # aliased source and destination are refused by the C ABI
        import ctypes as C
        hd = mesh.handle
        p = (C.c_void_p * 1)(dev[0].data_ptr())
        hd.check(hd.lib.scb_permute(hd.h, 50000, out[0].data_ptr(), 1, p, p, 1))


def _orders(n, keys, rng):
    srt = np.argsort(keys, kind="stable")
    degraded = srt.copy()
    # a fifth of the ordered bunch shuffled among itself: the order a tracking loop has after some steps
    pick = rng.choice(n, size=max(n // 5, 1), replace=False)
    degraded[pick] = degraded[rng.permutation(pick)]
    assert np.array_equal(np.sort(degraded), np.arange(n))
    return {"random": np.arange(n), "sorted": srt, "degraded": degraded, "reversed": srt[::-1].copy()}


@pytest.mark.parametrize("pdt,mdt,tol", [(np.float64, np.float64, TOL64), (np.float32, np.float32, TOL32),
                                         (np.float64, np.float32, TOL32), (np.float32, np.float64, TOL64)])
@pytest.mark.parametrize("n,grid", [(200003, (32, 32, 32)), (777, (9, 8, 7)), (60000, (64, 33, 17)), (20000, (200, 3, 2))])
@pytest.mark.parametrize("variant", ["cell", "cell_tile"])
def test_cell_order_kernels_match_the_oracle_for_any_order(scb, oracle, record, pdt, mdt, tol, n, grid, variant):
    import torch
    if grid[0] >= 200 and pdt == np.float32 and mdt == np.float32:
        # pure Float32 arithmetic: t = (x - lo) / delta reaches 199, so the fraction carries 199 * eps = 1.2e-5 and the
        # white-noise test field turns that into a 1.5e-5 difference from the Float64 oracle -- the reference's own
        # Float32 path has the same property; the long thin grid is there for the tile / key logic, covered in Float64
        pytest.skip("Float32 fractions on a 200-cell axis exceed the 1e-5 bar against the Float64 oracle by construction")
    x, y, z, q = gaussian(n, 42, dtype=pdt)
    mesh = scb.Mesh3D(grid, *to_dev(x, y, z), T=mdt)
    ref = oracle.mesh_from_particles(grid, x, y, z, T=np.float64 if tol == TOL32 else mdt)
    if tol == TOL32:
        ref.min_bounds, ref.max_bounds, ref.delta = (tuple(np.float64(v) for v in t)
                                                     for t in (mesh.min_bounds, mesh.max_bounds, mesh.delta))
    # Float64 meshes: the oracle runs on the particles' own element type (promotion as in the reference); Float32 meshes
    # are graded against the Float64 oracle on the same Float32-valued geometry
    xo, yo, zo, qo = (x, y, z, q) if tol == TOL64 else (a.astype(np.float64) for a in (x, y, z, q))
    oracle.deposit(ref, xo, yo, zo, qo, clamp=True)
    rng = np.random.default_rng(0)
    ref.efield[...] = rng.standard_normal(ref.efield.shape).astype(mdt)
    mesh.efield.copy_(torch.from_numpy(ref.efield.astype(mdt)).cuda())
    keys = oracle_keys(oracle, mesh, x, y, z)
    scb.set_particle_order(mesh, variant)   # "cell_tile": the shared-memory-tile deposit (kept for comparison)
    try:
        for name, order in _orders(n, keys, rng).items():
            want = oracle.interpolate_field(ref, xo[order], yo[order], zo[order], clamp=True)
            for shift in (0, 1):   # shift 1: arrays that start one element off a 32-byte boundary (element loads)
                dev = [torch.from_numpy(np.concatenate([np.zeros(shift, a.dtype), a[order]])).cuda()[shift:] for a in (x, y, z, q)]
                scb.deposit_(mesh, *dev)
                e = rel(mesh.rho.cpu().numpy(), ref.rho)
                record("rho %s shift %d" % (name, shift), e, tol)
                assert e < tol, (name, shift, e)
                got = scb.interpolate_field(mesh, *dev[:3])
                for g, w in zip(got, want):
                    if tol == TOL64:
                        # same weights, products and left-to-right sum as the reference: bit-identical
                        assert np.array_equal(g.cpu().numpy(), w), (name, shift)
                    else:
                        e = rel(g.cpu().numpy(), w)
                        record("interpolated %s shift %d" % (name, shift), e, tol)
                        assert e < tol, (name, shift, e)
        # clear=False accumulates (src/deposition.jl:218-247)
        dev = to_dev(*(a[np.argsort(keys, kind="stable")] for a in (x, y, z, q)))
        scb.deposit_(mesh, *dev)
        scb.deposit_(mesh, *dev, clear=False)
        assert rel(mesh.rho.cpu().numpy(), 2 * ref.rho) < tol
    finally:
        scb.set_particle_order(mesh, "random")


def test_cell_order_step_equals_default_step(scb):
    """the whole step with both kernel families on the same ordered bunch"""
    import torch
    x, y, z, q = gaussian(400000, 9)
    dev = to_dev(x, y, z, q)
    mesh = scb.Mesh3D((48, 40, 56), *dev[:3], gamma=3.0)
    _, sx, sy, sz, sq = scb.sort_particles_(mesh, *dev)
    outs_a = [torch.empty_like(sx) for _ in range(3)]
    outs_b = [torch.empty_like(sx) for _ in range(3)]
    scb.step_(mesh, sx, sy, sz, sq, *outs_a)
    rho_a = mesh.rho.clone()
    scb.set_particle_order(mesh, "cell")
    try:
        scb.step_(mesh, sx, sy, sz, sq, *outs_b)
    finally:
        scb.set_particle_order(mesh, "random")
    assert rel(mesh.rho.cpu().numpy(), rho_a.cpu().numpy()) < 1e-13
    for a, b in zip(outs_a, outs_b):
        assert rel(b.cpu().numpy(), a.cpu().numpy()) < 1e-12


def test_interpolate_kick_in_cell_order(scb):
    import torch
    x, y, z, q = gaussian(30000, 4)
    dev = to_dev(x, y, z, q)
    mesh = scb.Mesh3D((16, 16, 16), *dev[:3])
    scb.deposit_(mesh, *dev)
    scb.solve_(mesh)
    p0 = [torch.full_like(dev[0], v) for v in (1.0, 2.0, 3.0)]
    ex, ey, ez = scb.interpolate_field(mesh, *dev[:3])
    scb.set_particle_order(mesh, "cell")
    try:
        p = [t.clone() for t in p0]
        scb.interpolate_kick_(mesh, *dev[:3], *p, 0.25e-9, 1e-9)
    finally:
        scb.set_particle_order(mesh, "random")
    for got, base, e, c in zip(p, p0, (ex, ey, ez), (0.25e-9, 0.25e-9, 1e-9)):
        assert torch.allclose(got, base + c * e, rtol=1e-14, atol=0)


def test_auto_order_picks_the_kernel_family_and_keeps_the_result(scb):
    """SCB_ORDER_AUTO: same results as the explicit settings on a random and on an ordered bunch, and the ordered bunch
    must not fall into the default kernels' slow case (every lane of a warp reducing into one node)."""
    import torch
    x, y, z, q = gaussian(2_000_000, 12)
    dev = to_dev(x, y, z, q)
    mesh = scb.Mesh3D((64, 64, 64), *dev[:3])
    _, sx, sy, sz, sq = scb.sort_particles_(mesh, *dev)
    outs = [torch.empty_like(sx) for _ in range(3)]
    ref = {}
    for name, bunch in (("random", dev), ("ordered", (sx, sy, sz, sq))):
        scb.set_particle_order(mesh, "random")
        scb.step_(mesh, *bunch, *outs)
        ref[name] = (mesh.rho.clone(), [o.clone() for o in outs])
    scb.set_particle_order(mesh, "auto")
    try:
        for name, bunch in (("random", dev), ("ordered", (sx, sy, sz, sq)), ("random", dev)):
            for _ in range(3):
                scb.step_(mesh, *bunch, *outs)
            assert rel(mesh.rho.cpu().numpy(), ref[name][0].cpu().numpy()) < 1e-13
            for a, b in zip(outs, ref[name][1]):
                assert rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-12
        # the choice shows in the deposit time of the ordered bunch: run-accumulating kernels, not 8 colliding reductions
        hd = mesh.handle
        hd.enable_timing(True)
        t = {}
        for name, order in (("auto", "auto"), ("random", "random"), ("cell", "cell")):
            scb.set_particle_order(mesh, order)
            best = 1e9
            for _ in range(4):
                scb.step_(mesh, sx, sy, sz, sq, *outs)
                best = min(best, hd.timing()["deposit_ms"])
            t[name] = best
        hd.enable_timing(False)
        if t["random"] > 2.0 * t["cell"]:   # the two families are separated by more than the timer's noise
            assert t["auto"] < 0.5 * (t["random"] + t["cell"]), t
    finally:
        scb.set_particle_order(mesh, "random")
