"""Two-GPU parity of the particle-sharded step: slab-decomposed solve (reduce-scatter, all-to-all
pencil transposes, all-gather over NCCL) against the single-GPU result on the same particles.
Needs two B200s (`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    from __graft_entry__ import load_package
    scb = load_package()
    from spacecharge_jl_b200.sharding import shard_range

    worst = 0.0
    for grid, cath, T, tol in (((32, 24, 16), False, np.float64, 1e-12), ((16, 20, 32), True, np.float64, 1e-11),
                               ((32, 32, 32), False, np.float32, 2e-6)):
        rng = np.random.default_rng(42)
        n = 200001
        x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
        if cath:
            z = z + 6e-3
        q = np.full(n, 1e-9 / n)
        dev = "cuda:%d" % rank
        full = [torch.from_numpy(a).to(dev) for a in (x, y, z, q)]
        b, e = shard_range(n, rank, world)
        mine = [t[b:e].contiguous() for t in full]
        # single-GPU result on all particles (every rank computes it for itself)
        ref = scb.Mesh3D(grid, *full[:3], T=T, gamma=2.0)
        scb.deposit_(ref, *full)
        scb.solve_(ref, at_cathode=cath)
        rex = scb.interpolate_field(ref, *mine[:3])
        for sharded in (True, False):
            mesh = scb.Mesh3D(grid, *mine[:3], T=T, gamma=2.0, group=dist.group.WORLD, sharded_solve=sharded)
            assert mesh.sharded == sharded
            assert mesh.min_bounds == ref.min_bounds and mesh.delta == ref.delta
            scb.deposit_(mesh, *mine)
            scb.solve_(mesh, at_cathode=cath)
            out = scb.interpolate_field(mesh, *mine[:3])
            torch.cuda.synchronize()
            for c in range(3):
                a, r = mesh.efield[..., c], ref.efield[..., c]
                err = float((a - r).abs().max() / r.abs().max())
                worst = max(worst, err)
                assert err < tol, (grid, cath, sharded, c, err)
                err = float((out[c] - rex[c]).abs().max() / rex[c].abs().max())
                assert err < tol, (grid, cath, sharded, "interp", c, err)
            # fused step; in the sharded mode also with the field slabs broadcast and gathered in overlapping passes
            for overlap in ("0", "1"):
                os.environ["SCB_GATHER_OVERLAP"] = overlap
                outs = [torch.full_like(mine[0], float("nan")) for _ in range(3)]
                mesh.efield.zero_()
                scb.step_(mesh, *mine, *outs, at_cathode=cath)
                torch.cuda.synchronize()
                for c in range(3):
                    err = float((outs[c] - rex[c]).abs().max() / rex[c].abs().max())
                    assert err < tol, (grid, cath, sharded, overlap, "step interp", c, err)
                    err = float((mesh.efield[..., c] - ref.efield[..., c]).abs().max() / ref.efield[..., c].abs().max())
                    assert err < tol, (grid, cath, sharded, overlap, "step field", c, err)
            os.environ.pop("SCB_GATHER_OVERLAP", None)
            # host-buffer shards (scb_step_host_sharded_async): one blocking step, then three queued back to back
            hin = [t.cpu().pin_memory() for t in mine]
            houts = [[torch.full_like(hin[0], float("nan")).pin_memory() for _ in range(3)] for _ in range(3)]
            mesh.efield.zero_()
            scb.step_host_(mesh, *hin, *houts[0], at_cathode=cath)
            for k in range(3):
                scb.step_host_async_(mesh, *hin, *houts[k], at_cathode=cath)
            scb.step_host_wait_(mesh)
            for k in range(3):
                for c in range(3):
                    err = float((houts[k][c].to(dev) - rex[c]).abs().max() / rex[c].abs().max())
                    assert err < tol, (grid, cath, sharded, "host step interp", k, c, err)
            for c in range(3):
                err = float((mesh.efield[..., c] - ref.efield[..., c]).abs().max() / ref.efield[..., c].abs().max())
                assert err < tol, (grid, cath, sharded, "host step field", c, err)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("%g" % worst)
    dist.destroy_process_group()


def test_sharded_step_matches_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
