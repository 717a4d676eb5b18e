"""Multi-GPU parity of the particle-sharded step at world sizes 2, 4 and 8: slab-decomposed solve (reduce-scatter,
pencil transposes over peer memory / NCCL, all-gather) and the replicated-solve fallback, against the single-GPU
result on the same particles AND, for the BASELINE-sized grids (128^3 free space, 128x128x256 cathode), element-wise
against the oracle's C restatement (oracle/cpu_reference.py, computed once by the parent process).
Needs `gpurun --gpus N`; every world size larger than the box's GPU count is skipped (the driver's one-GPU GPUTEST
skips all of them: profiles/r02_pytest_multi8.log is the committed 8-GPU run)."""
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
    for grid, cath, T, tol, n in CASES:
        x, y, z, q = _case_particles(n, cath)
        dev = "cuda:%d" % rank
        full = [torch.from_numpy(a).to(dev) for a in (x, y, z, q)]
        b, e = shard_range(n, rank, world)
        mine = [t[b:e].contiguous() for t in full]
        # single-GPU result on all particles (every rank computes it for itself)
        ref = scb.Mesh3D(grid, *full[:3], T=T, gamma=2.0)
        scb.deposit_(ref, *full)
        scb.solve_(ref, at_cathode=cath)
        rex = scb.interpolate_field(ref, *mine[:3])
        oracle_file = os.path.join(out_dir, "oracle_%dx%dx%d_%d.npz" % (grid + (int(cath),)))
        oracle_e = np.load(oracle_file)["efield"] if os.path.exists(oracle_file) else None
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
                if oracle_e is not None:   # element-wise against the C restatement of the reference, 1e-10
                    w = torch.from_numpy(oracle_e[..., c]).to(dev)
                    err = float((a - w).abs().max() / w.abs().max())
                    assert err < 1e-10, (grid, cath, sharded, "oracle", c, err)
                    if rank == 0:
                        print("world %d grid %s cathode %d sharded %d: E%d vs oracle %.2e, vs one GPU %.2e"
                              % (world, grid, cath, sharded, c, err, float((a - r).abs().max() / r.abs().max())), flush=True)
            if grid[0] <= 32 and T == np.float64:
                # scalar potential on a particle-sharded mesh (extension): equals the single-GPU potential
                scb.deposit_(ref, *full)
                scb.solve_potential_(ref, at_cathode=cath)
                scb.deposit_(mesh, *mine)
                scb.solve_potential_(mesh, at_cathode=cath)
                torch.cuda.synchronize()
                err = float((mesh.phi - ref.phi).abs().max() / ref.phi.abs().max())
                assert err < 1e-11, (grid, cath, sharded, "phi", err)
            if not sharded and grid[0] <= 32:
                # two deposits into one grid (clear=False, e.g. two species): in the replicated mode rho holds the sum
                # over the ranks after every call, so only the SECOND call's contribution may be reduced again
                half = mine[0].numel() // 2
                scb.deposit_(mesh, *[t[:half] for t in mine])
                scb.deposit_(mesh, *[t[half:] for t in mine], clear=False)
                torch.cuda.synchronize()
                err = float((mesh.rho - ref.rho).abs().max() / ref.rho.abs().max())
                assert err < max(tol, 1e-13), (grid, "two-species deposit with clear=False over ranks", err)
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


# (grid, cathode, mesh type, tolerance against the single-GPU result, particles)
CASES = (((32, 24, 16), False, np.float64, 1e-12, 200001), ((16, 20, 32), True, np.float64, 1e-11, 200001),
         ((32, 32, 32), False, np.float32, 2e-6, 200001),
         ((128, 128, 128), False, np.float64, 1e-12, 2000003), ((128, 128, 256), True, np.float64, 1e-11, 2000003))


def _case_particles(n, cath):
    rng = np.random.default_rng(42)
    x, y, z = (rng.standard_normal(n) * 1e-3 for _ in range(3))
    if cath:
        z = z + 6e-3
    return x, y, z, np.full(n, 1e-9 / n)


def _write_oracle_fields(out_dir):
    """E of the BASELINE-sized cases from the C restatement of the reference, once, for every worker to compare with"""
    sys.path.insert(0, ROOT)
    from oracle import spacecharge_oracle as so
    from oracle.cpu_reference import RefPort
    for grid, cath, T, _, n in CASES:
        if grid[0] < 128 or T != np.float64:
            continue
        x, y, z, q = _case_particles(n, cath)
        m = so.mesh_from_particles(grid, x, y, z, gamma=2.0)
        rp = RefPort(grid, m.min_bounds, m.delta, 2.0)
        rp.deposit(x, y, z, q)
        rp.solve(cath, m.max_bounds)
        np.savez(os.path.join(out_dir, "oracle_%dx%dx%d_%d.npz" % (grid + (int(cath),))), efield=rp.efield)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_matches_single_gpu_and_oracle(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _write_oracle_fields(str(tmp_path))
    port = 29600 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
