"""The secondary GPU baseline (reference structure restated in plain CUDA + cuFFT,
baseline/naive_gpu/) must itself agree with the oracle, otherwise timing it means nothing."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("at_cathode", [False, True])
def test_naive_gpu_baseline_matches_oracle(oracle, at_cathode):
    import torch
    spec = importlib.util.spec_from_file_location("naive_gpu_driver", os.path.join(ROOT, "baseline", "naive_gpu", "driver.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    rng = np.random.default_rng(42)
    n = 100000
    x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
    z = z + 6e-3
    q = np.full(n, 1e-9 / n)
    grid = (16, 12, 20)
    ref, want = oracle.full_step(grid, x, y, z, q, gamma=2.0, at_cathode=at_cathode)
    d = [torch.from_numpy(a).cuda() for a in (x, y, z, q)]
    rho = torch.zeros(grid[::-1], dtype=torch.float64, device="cuda")
    e = torch.zeros((3,) + grid[::-1], dtype=torch.float64, device="cuda")
    out = [torch.empty_like(d[0]) for _ in range(3)]
    drv.RefGpu().step(grid, ref.min_bounds, ref.max_bounds, ref.delta, 2.0, at_cathode, *d, rho, e, *out)
    torch.cuda.synchronize()
    got_rho = rho.permute(2, 1, 0).cpu().numpy()
    got_e = e.permute(3, 2, 1, 0).cpu().numpy()
    assert np.abs(got_rho - ref.rho).max() / np.abs(ref.rho).max() < 1e-12
    for c in range(3):
        assert np.abs(got_e[..., c] - ref.efield[..., c]).max() / np.abs(ref.efield[..., c]).max() < 1e-10
        assert np.abs(out[c].cpu().numpy() - want[c]).max() / np.abs(want[c]).max() < 1e-10
