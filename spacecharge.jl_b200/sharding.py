"""Particle sharding for multi-GPU runs (one process per GPU, torch.distributed).

The reference is single-device; this is the new part of section 8(e): particles are
split evenly in index order, every rank deposits its shard into a full-size private rho, the grids
are summed over the ranks, every rank solves and interpolates its own particles.  The helpers here
are backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Sequence, Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the particles owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def global_extrema(lo: Sequence[float], hi: Sequence[float], group, device="cpu"):
    """Element-wise min of `lo` and max of `hi` over the ranks of `group` (Float64 carries both
    Float32 and Float64 extrema exactly)."""
    import torch
    import torch.distributed as dist
    tlo = torch.tensor([float(v) for v in lo], dtype=torch.float64, device=device)
    thi = torch.tensor([float(v) for v in hi], dtype=torch.float64, device=device)
    dist.all_reduce(tlo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(thi, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in tlo.tolist()], [float(v) for v in thi.tolist()]


def allreduce_rho(rho, group) -> None:
    """Sum the per-rank charge grids in place."""
    import torch.distributed as dist
    dist.all_reduce(rho, op=dist.ReduceOp.SUM, group=group)
