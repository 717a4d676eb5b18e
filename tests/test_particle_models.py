"""CPU checks of the particle-kernel layouts (tests/particle_model.py) against the oracle: the cell-tile accumulator
with its fold reproduces the reference's deposit, the node-major packed records with the lane-pair summation reproduce
the reference's interpolation (Float64: to the last ulps; Float32 records, thread per particle: bit for bit)."""
import numpy as np
import pytest

import particle_model as pm


def bunch(n, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    x, y, z = ((rng.standard_normal(n) * 1e-3).astype(dtype) for _ in range(3))
    q = (rng.uniform(0.5, 1.5, n) * 1e-9 / n).astype(dtype)
    return x, y, z, q


@pytest.mark.parametrize("grid", [(8, 8, 8), (5, 9, 12), (2, 2, 2), (16, 3, 7)])
def test_tile_deposit_and_fold_equal_the_reference_deposit(oracle, grid):
    x, y, z, q = bunch(4000, sum(grid))
    mesh = oracle.mesh_from_particles(grid, x, y, z)
    oracle.deposit(mesh, x, y, z, q, clamp=True)
    tiles, rho = pm.deposit_tiles(grid, mesh.min_bounds, mesh.delta, x, y, z, q)
    assert np.max(np.abs(rho - mesh.rho)) <= 1e-14 * np.max(np.abs(mesh.rho))
    # every contribution lands in exactly one tile slot: the accumulator conserves the charge by itself
    assert abs(tiles.sum() - q.sum()) <= 1e-13 * abs(q.sum())
    # the slots of the last cell row / column / plane are never addressed (cell index clamped to n-2)
    assert not tiles[-1, :, :, :].any() and not tiles[:, -1, :, :].any()


def test_tile_deposit_float32_mesh(oracle):
    grid = (6, 7, 8)
    x, y, z, q = bunch(3000, 5, np.float32)
    mesh = oracle.mesh_from_particles(grid, x, y, z, T=np.float32)
    oracle.deposit(mesh, x, y, z, q, clamp=True)
    _, rho = pm.deposit_tiles(grid, mesh.min_bounds, mesh.delta, x, y, z, q, T=np.float32)
    assert np.max(np.abs(rho - mesh.rho)) <= 2e-5 * np.max(np.abs(mesh.rho))    # Float32 summation order differs


@pytest.mark.parametrize("grid", [(8, 8, 8), (5, 9, 12), (2, 2, 2)])
def test_lane_pair_gather_equals_the_reference_interpolation(oracle, grid):
    x, y, z, _ = bunch(5000, 3 + sum(grid))
    mesh = oracle.mesh_from_particles(grid, x, y, z)
    rng = np.random.default_rng(1)
    mesh.efield[...] = rng.standard_normal(mesh.efield.shape)
    want = oracle.interpolate_field(mesh, x, y, z, clamp=True)
    got = pm.gather_pair_f64(pm.pack_f64(mesh.efield), mesh.min_bounds, mesh.delta, x, y, z)
    scale = np.max(np.abs(mesh.efield))
    for a, b in zip(got, want):
        assert np.max(np.abs(a - b)) <= 8 * np.finfo(np.float64).eps * scale     # x halves summed separately


@pytest.mark.parametrize("pdt", [np.float32, np.float64])
def test_packed_float32_gather_is_bit_identical(oracle, pdt):
    grid = (9, 6, 11)
    x, y, z, _ = bunch(4000, 17, pdt)
    mesh = oracle.mesh_from_particles(grid, x, y, z, T=np.float32)
    rng = np.random.default_rng(2)
    mesh.efield[...] = rng.standard_normal(mesh.efield.shape).astype(np.float32)
    want = oracle.interpolate_field(mesh, x, y, z, clamp=True)
    got = pm.gather_packed_f32(pm.pack_f32(mesh.efield), mesh.min_bounds, mesh.delta, x, y, z)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and np.array_equal(a, b)


# ---- cell-ordered regime (csrc/sorted.cu) -----------------------------------------------------------------------------
@pytest.mark.parametrize("n,grid", [(1, (4, 4, 4)), (5000, (8, 8, 8)), (9000, (16, 5, 3)), (4097, (40, 2, 2))])
def test_tile_ranked_radix_sort_is_the_stable_sort_by_cell_key(oracle, n, grid):
    x, y, z, q = bunch(n, 7 + n)
    mesh = oracle.mesh_from_particles(grid, x, y, z)
    keys = pm.cell_keys(grid, mesh.min_bounds, mesh.delta, x, y, z)
    bits = max(1, int(np.ceil(np.log2(grid[0] * grid[1] * grid[2]))))
    perm = pm.radix_sort_pairs(keys, bits, tile=1024, warp_keys=128)
    assert np.array_equal(perm, np.argsort(keys, kind="stable"))


@pytest.mark.parametrize("order", ["random", "sorted", "drifted"])
def test_run_deposit_with_lookahead_and_lane_scan_equals_the_reference_deposit(oracle, order):
    grid = (8, 6, 7)
    x, y, z, q = bunch(6000, 99)
    mesh = oracle.mesh_from_particles(grid, x, y, z)
    keys = pm.cell_keys(grid, mesh.min_bounds, mesh.delta, x, y, z)
    if order != "random":
        p = np.argsort(keys, kind="stable")
        if order == "drifted":   # a fifth of the ordered bunch shuffled among itself
            rng = np.random.default_rng(5)
            pick = rng.choice(len(p), len(p) // 5, replace=False)
            p[pick] = p[rng.permutation(pick)]
        x, y, z, q = x[p], y[p], z[p], q[p]
    oracle.deposit(mesh, x, y, z, q, clamp=True)
    rho, flushes = pm.deposit_runs(grid, mesh.min_bounds, mesh.delta, x, y, z, q)
    assert np.max(np.abs(rho - mesh.rho)) <= 1e-14 * np.max(np.abs(mesh.rho))
    assert abs(rho.sum() - q.sum()) <= 1e-13 * abs(q.sum())
    # what the regime is about: an ordered bunch reaches memory far less often than once per particle
    if order == "sorted":
        assert flushes < 0.15 * len(x)
    if order == "random":
        assert flushes > 0.8 * len(x)
