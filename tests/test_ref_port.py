"""The C restatement used for the timed CPU baseline agrees with the NumPy oracle."""
import numpy as np
import pytest

from oracle import spacecharge_oracle as so
from oracle.cpu_reference import RefPort


@pytest.mark.parametrize("at_cathode", [False, True])
def test_c_port_matches_numpy_oracle(at_cathode):
    rng = np.random.default_rng(42)
    n = 50000
    x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
    z = z + 6e-3
    q = np.full(n, 1e-9 / n)
    grid = (12, 16, 20)
    ref, want = so.full_step(grid, x, y, z, q, gamma=2.0, at_cathode=at_cathode)
    rp = RefPort(grid, ref.min_bounds, ref.delta, gamma=2.0)
    got, _ = rp.timed_step(x, y, z, q, at_cathode, ref.max_bounds)
    assert np.array_equal(rp.rho, ref.rho)  # same serial order, same arithmetic: bit-exact
    for c in range(3):
        assert np.abs(rp.efield[..., c] - ref.efield[..., c]).max() / np.abs(ref.efield[..., c]).max() < 1e-11
        assert np.abs(got[c] - want[c]).max() / np.abs(want[c]).max() < 1e-11
