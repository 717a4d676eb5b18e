"""Host-side contract of the kx-slab multi-GPU solve, on a CPU box: the buffer layouts and the index arithmetic of the two
fused exchanges (tests/slab_model.py restates them from csrc/api.cu / csrc/fft_passes.cuh) reproduce the single-rank
solve for 2, 4 and 8 model ranks, including ranks that own no valid kx at all."""
import numpy as np
import pytest

import fused_model as fm
import slab_model as sm


@pytest.mark.parametrize("grid,G", [((6, 5, 8), 2), ((6, 5, 8), 4), ((5, 4, 8), 8), ((12, 3, 4), 4), ((3, 3, 16), 8)])
def test_kx_slab_layouts_reproduce_the_single_rank_solve(oracle, grid, G):
    rng = np.random.default_rng(sum(grid) + G)
    rho = rng.standard_normal(grid)
    delta, gamma = (1.1e-4, 0.9e-4, 1.3e-4), 1.7
    want = fm.solve_fused(rho, delta, gamma)
    got = sm.solve_kx_slabs(rho, delta, gamma, G)
    for c in range(3):
        assert np.abs(got[..., c] - want[..., c]).max() <= 1e-12 * np.abs(want[..., c]).max()
    # and the single-rank model itself is the oracle's reference-structured solve
    mesh = oracle.mesh_from_bounds(grid, (0, 0, 0), tuple(d * (n - 1) for d, n in zip(delta, grid)), gamma=gamma)
    mesh.delta = tuple(np.float64(d) for d in delta)
    mesh.rho[...] = rho
    oracle.solve(mesh)
    for c in range(3):
        assert np.abs(got[..., c] - mesh.efield[..., c]).max() <= 1e-10 * np.abs(mesh.efield[..., c]).max()
