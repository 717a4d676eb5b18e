"""World-size-2 gloo test of the particle-sharded path's host logic (section 8e): global bounds,
shard ranges and the sum of the per-rank charge grids.  The per-rank deposit is done by the oracle
here (no GPU); on the GPU box the same plumbing runs over NCCL with the CUDA deposit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    from oracle import spacecharge_oracle as so
    scb = load_package()
    from spacecharge_jl_b200.sharding import allreduce_rho, shard_range

    rng = np.random.default_rng(42)
    x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
    q = np.full(n, 1e-9 / n)
    b, e = shard_range(n, rank, world)
    grid = (16, 12, 10)
    lo, hi, d = scb.Mesh3D._auto_bounds(grid, x[b:e], y[b:e], z[b:e], np.float64, 0, None, dist.group.WORLD)
    whole = so.mesh_from_particles(grid, x, y, z)
    assert lo == whole.min_bounds and hi == whole.max_bounds and d == whole.delta
    local = so.mesh_from_bounds(grid, lo, hi)
    local.min_bounds, local.max_bounds, local.delta = lo, hi, d
    so.deposit(local, x[b:e], y[b:e], z[b:e], q[b:e])
    t = torch.from_numpy(np.ascontiguousarray(local.rho.transpose(2, 1, 0)))
    allreduce_rho(t, dist.group.WORLD)
    so.deposit(whole, x, y, z, q)
    got = t.numpy().transpose(2, 1, 0)
    err = np.abs(got - whole.rho).max() / np.abs(whole.rho).max()
    assert err < 1e-13, err
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("%g" % err)
    dist.destroy_process_group()


def test_sharded_deposit_matches_single_process(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, 20001, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
