"""Host-side mirror logic that runs without a GPU: constructor arithmetic and error behaviour."""
import numpy as np
import pytest

from oracle import spacecharge_oracle as so


@pytest.mark.parametrize("pdt", [np.float32, np.float64])
@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_auto_bounds_arithmetic_is_bit_identical_to_the_oracle(scb, pdt, T):
    rng = np.random.default_rng(5)
    x, y, z = ((rng.standard_normal(1000) * s).astype(pdt) for s in (1e-3, 2e-3, 0.5e-3))
    grid = (16, 33, 8)
    lo, hi, d = scb.Mesh3D._auto_bounds(grid, x, y, z, np.dtype(T).type, 0, None)
    rlo, rhi, rd = so.auto_bounds(grid, x, y, z, T)
    assert lo == rlo and hi == rhi and d == rd
    assert all(type(v) is np.dtype(T).type for v in lo + hi + d)


def test_zero_extent_gets_the_1e_6_spacing(scb):
    lo, hi, d = scb.Mesh3D._auto_bounds((4, 4, 4), [0.0], [0.0], [0.0], np.float64, 0, None)
    assert d == (1e-6, 1e-6, 1e-6)                      # src/mesh.jl:145-149


def test_constructor_errors_before_touching_the_device(scb):
    E = scb.ErrorException
    for bad in ((0, 2, 2), (2, 1, 2), (2, 2, 1)):
        with pytest.raises(E):
            scb.Mesh3D(bad, [0.0], [0.0], [0.0])
    with pytest.raises(E):
        scb.Mesh3D((2, 2, 2), [], [], [])
    with pytest.raises(E):
        scb.Mesh3D((4, 4, 4), [0.0, 1.0], [0.0], [0.0])
    with pytest.raises(E):
        scb.Mesh3D((2, 2, 2), (0, 0, 0), (0, 0, 0))
    with pytest.raises(E):
        scb.Mesh3D((2, 2, 2), (1, 1, 1), (0, 0, 0))


def test_shard_ranges_cover_everything(scb):
    from spacecharge_jl_b200.sharding import shard_range
    for n, w in ((10, 3), (100_000_000, 8), (7, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def test_bind_host_to_device_is_harmless_without_a_gpu(scb):
    """bind_host_to_device is host plumbing: without NVML / a GPU it returns None and leaves the affinity alone; with
    SCB_NUMA_BIND=0 it never touches it."""
    import os
    if not hasattr(os, "sched_getaffinity"):
        return
    before = os.sched_getaffinity(0)
    os.environ["SCB_NUMA_BIND"] = "0"
    try:
        assert scb.bind_host_to_device(0) is None
    finally:
        os.environ.pop("SCB_NUMA_BIND")
    out = scb.bind_host_to_device(0)
    try:
        assert out is None or set(out) <= before
    finally:
        os.sched_setaffinity(0, before)
