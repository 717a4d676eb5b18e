"""spacecharge.jl_b200 -- host-side mirror of SpaceCharge.jl's public surface over the C ABI.

The reference's whole API is five names (src/SpaceCharge.jl:17): ``Mesh3D, deposit!,
clear_mesh!, interpolate_field, solve!``.  Julia is not installed in this image, so the tested
host side is this Python twin of the Julia shim in ``julia/SpaceChargeB200.jl``: same names
(``!`` spelled as a trailing underscore), same argument meaning, same error behaviour
(``ErrorException`` with the reference's messages).  All computation happens in
``lib/libspacecharge_b200.so`` (hand-written sm_100a CUDA); torch only provides device memory,
streams and torch.distributed.  There is no CPU fallback.

Arrays follow the reference's column-major layout: ``mesh.rho[ix, iy, iz]`` and
``mesh.efield[ix, iy, iz, c]`` are torch views whose x index is the fastest in memory.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import LibraryMissing, ScbError  # noqa: F401

__all__ = ["Mesh3D", "deposit_", "clear_mesh_", "interpolate_field", "solve_", "solve_freespace_",
           "get_green_function_", "cell_indices", "step_", "step_host_", "step_host_async_", "step_host_wait_", "ErrorException", "CLIGHT", "FPEI",
           "solve_potential_", "magnetic_field", "interpolate_kick_",
           "sort_particles", "permute", "sort_particles_", "set_particle_order", "particle_order_fraction",
           "Handle", "default_handle", "bind_host_to_device"]

CLIGHT = 299792458.0          # src/utils.jl:7
FPEI = CLIGHT ** 2 * 1.0e-7   # src/utils.jl:8


class ErrorException(Exception):
    """Julia's ``error("...")`` (src/mesh.jl:106-116, 206-211; src/deposition.jl:226-228)."""


def _torch():
    import torch
    return torch


def _np_dtype(T):
    torch = _torch()
    if T in (np.float32, "float32", "Float32", torch.float32):
        return np.float32
    if T in (np.float64, "float64", "Float64", float, torch.float64):
        return np.float64
    raise ErrorException("T must be Float32 or Float64")


def _torch_dtype(npdt):
    torch = _torch()
    return torch.float32 if np.dtype(npdt) == np.float32 else torch.float64


def _tag(torch_dtype):
    torch = _torch()
    if torch_dtype == torch.float32:
        return _lib.SCB_F32
    if torch_dtype == torch.float64:
        return _lib.SCB_F64
    raise ErrorException("only Float32 and Float64 arrays are supported")


# ------------------------------------------------------------------------------------ handle
class Handle:
    """One ``scb_handle`` per (device, stream-in-use).  Not thread-safe (same as the C ABI)."""

    def __init__(self, device: int = 0, green_cache: bool = True):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise LibraryMissing("no CUDA device: spacecharge.jl_b200 has no CPU fallback")
        self.device = int(device)
        self._stream = torch.cuda.current_stream(self.device).cuda_stream
        opt = _lib.scb_options()
        opt.green_cache = 1 if green_cache else 0
        h = C.c_void_p()
        rc = self.lib.scb_create(self.device, C.c_void_p(self._stream), C.byref(opt), C.byref(h))
        if rc != 0:
            raise ScbError(rc, "scb_create failed (an sm_100 GPU is required)")
        self.h = h

    def check(self, rc):
        if rc != 0:
            raise ScbError(rc, self.lib.scb_last_error(self.h).decode())

    def use_current_stream(self):
        s = _torch().cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            self.check(self.lib.scb_set_stream(self.h, C.c_void_p(s)))
            self._stream = s

    def sync(self):
        self.check(self.lib.scb_sync(self.h))

    def enable_timing(self, on=True):
        self.check(self.lib.scb_enable_timing(self.h, 1 if on else 0))

    def timing(self) -> dict:
        t = _lib.scb_timing()
        self.check(self.lib.scb_get_timing(self.h, C.byref(t)))
        return {"deposit_ms": t.deposit_ms, "solve_ms": t.solve_ms, "interpolate_ms": t.interpolate_ms,
                "green_ms": t.green_ms, "pass_ms": list(t.pass_ms)[:5],
                # slab-decomposed solve: the collectives inside pass_ms[0] (F1) and pass_ms[4] (B3), timed on their own
                "reduce_scatter_ms": t.pass_ms[5], "all_gather_ms": t.pass_ms[6]}

    def launch_count(self) -> int:
        return int(self.lib.scb_launch_count(self.h))

    def workspace_bytes(self) -> int:
        return int(self.lib.scb_workspace_bytes(self.h))

    def drop_green_cache(self):
        self.check(self.lib.scb_drop_green_cache(self.h))

    def set_particle_order(self, order):
        """``"random"`` (default), ``"cell"``, ``"cell_tile"`` or ``"auto"`` (the handle samples the bunch's order every
        eighth deposit and picks): which kernels the particle passes use (scb_set_particle_order)."""
        code = {"random": _lib.SCB_ORDER_RANDOM, "cell": _lib.SCB_ORDER_CELL, "cell_tile": _lib.SCB_ORDER_CELL_TILE,
                "auto": _lib.SCB_ORDER_AUTO}.get(order, order)
        self.check(self.lib.scb_set_particle_order(self.h, int(code)))

    def init_comm(self, group):
        """Create the library's own NCCL communicator for `group` (one rank per GPU): rank 0 draws
        the unique id, torch.distributed broadcasts it."""
        import torch
        import torch.distributed as dist
        if getattr(self, "_comm_group", None) is group:
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            rc = self.lib.scb_comm_unique_id(buf)
            if rc != 0:
                raise ScbError(rc, "scb_comm_unique_id failed (libnccl.so.2 not loadable?)")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        dev = "cuda:%d" % self.device if dist.get_backend(group) == "nccl" else "cpu"
        uid = uid.to(dev)
        dist.broadcast(uid, src=dist.get_global_rank(group, 0), group=group)
        raw = (C.c_ubyte * 128)(*uid.cpu().tolist())
        self.check(self.lib.scb_comm_init(self.h, world, rank, raw))
        self._comm_group = group

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.scb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bind_host_to_device(device: Optional[int] = None) -> Optional[list]:
    """Pin the calling process to the CPU cores NVML reports as local to `device` (its socket / NUMA node), so that
    pinned host buffers allocated afterwards are placed next to the GPU's PCIe root port (first-touch policy) and the
    host-buffer steps (scb_step_host*) do not cross the inter-socket link.  Call it before allocating the buffers.
    Returns the CPU list that was set, or None when NVML, the affinity query or sched_setaffinity is unavailable or
    SCB_NUMA_BIND=0.  Host plumbing only: no effect on any device result."""
    import os
    if os.environ.get("SCB_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    torch = _torch()
    try:
        import pynvml
        if device is None:
            device = torch.cuda.current_device()
        prop = torch.cuda.get_device_properties(device)
        pynvml.nvmlInit()
        bus = "%08x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        allowed = sorted(local & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


_handles = {}


def default_handle(device: Optional[int] = None) -> Handle:
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]


# ------------------------------------------------------------------------------------- mesh
def _device_array(a, device):
    """Particle array -> 1-D contiguous CUDA tensor (the reference takes CuArrays; host arrays
    are uploaded like ``CuArray(x)``)."""
    torch = _torch()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype.kind != "f" or arr.dtype.itemsize not in (4, 8):
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if t.device.type != "cuda":
        t = t.to("cuda:%d" % device)
    return t.contiguous().view(-1)


def _particle_view(a, device, allow_broadcast=False):
    """Particle array -> (CUDA tensor, element stride) WITHOUT copying when ``a`` is a 1-D CUDA tensor
    view with a positive stride -- e.g. ``records[:, 0]`` of an (Np, 6) phase-space array, the layout of
    Bmad-style callers.  ``allow_broadcast``: a stride-0 view (``q0.expand(Np)``) means
    one charge for every particle.  Anything else goes through ``_device_array`` (dense copy, stride 1)."""
    torch = _torch()
    if (isinstance(a, torch.Tensor) and a.device.type == "cuda" and a.dim() == 1 and a.numel() > 0
            and a.dtype in (torch.float32, torch.float64)):
        st = a.stride(0)
        if st >= 1 or (allow_broadcast and st == 0):
            return a, int(st)
    return _device_array(a, device), 1


def _strides(x=1, y=1, z=1, q=1, ex=1, ey=1, ez=1):
    return _lib.scb_particle_strides(x, y, z, q, ex, ey, ez, 0)


def _extrema_host(a):
    arr = np.asarray(a)
    if arr.dtype.kind != "f" or arr.dtype.itemsize not in (4, 8):
        arr = arr.astype(np.float64)
    return arr.dtype.type(arr.min()), arr.dtype.type(arr.max()), arr.dtype.type


class Mesh3D:
    """src/mesh.jl:19-34.  Two constructors, selected like Julia's dispatch:

    ``Mesh3D(grid_size, particles_x, particles_y, particles_z; T, gamma, total_charge)``
        bounds from the particle extrema (src/mesh.jl:95-174)
    ``Mesh3D(grid_size, min_bounds, max_bounds; T, gamma, total_charge)``
        manual bounds (src/mesh.jl:196-238)
    """

    def __init__(self, grid_size: Sequence[int], *args, T=np.float64, gamma: float = 1.0,
                 total_charge: float = 0.0, device: Optional[int] = None, handle: Optional[Handle] = None,
                 group=None, sharded_solve: Optional[bool] = None):
        torch = _torch()
        grid_size = tuple(int(g) for g in grid_size)
        if len(grid_size) != 3:
            raise ErrorException("grid_size must have three elements")
        npdt = _np_dtype(T)
        Tn = np.dtype(npdt).type
        if any(g <= 1 for g in grid_size):
            raise ErrorException("All elements of grid_size must be at least 2.")
        if len(args) == 2:
            lo, hi = args
            if any(h <= l for h, l in zip(hi, lo)):
                raise ErrorException("max_bounds must be strictly greater than min_bounds for all dimensions.")
            lo = tuple(Tn(v) for v in lo)
            hi = tuple(Tn(v) for v in hi)
            delta = tuple(Tn((h - l) / Tn(n - 1)) for h, l, n in zip(hi, lo, grid_size))  # :214-220
            dev_hint = None
        elif len(args) == 3:
            px, py, pz = args
            if len(px) == 0 or len(py) == 0 or len(pz) == 0:
                raise ErrorException("Particle arrays cannot be empty.")
            if not (len(px) == len(py) == len(pz)):
                raise ErrorException("Particle coordinate arrays must have the same length.")
            dev_hint = px.device.index if isinstance(px, torch.Tensor) and px.device.type == "cuda" else None
            lo, hi, delta = self._auto_bounds(grid_size, px, py, pz, Tn, dev_hint if device is None else device, handle, group)
        else:
            raise ErrorException("Mesh3D(grid_size, x, y, z) or Mesh3D(grid_size, min_bounds, max_bounds)")
        if device is None:
            device = dev_hint if dev_hint is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.grid_size = grid_size
        self.min_bounds = lo
        self.max_bounds = hi
        self.delta = delta
        self.gamma = Tn(gamma)
        self.total_charge = Tn(total_charge)
        self.T = npdt
        self.device = int(device)
        self.handle = handle if handle is not None else default_handle(self.device)
        self.group = group  # torch.distributed process group for particle-sharded runs (or None)
        # slab-decomposed solve (scb_solve_sharded) when the grid divides over the ranks; otherwise
        # rho is all-reduced and the solve replicated.  In the sharded mode mesh.rho holds this
        # rank's PARTIAL charge grid after deposit_ (call reduce_rho_() to materialise the sum).
        # sharded_solve=None: the slab-decomposed solve whenever the grid divides over the ranks.  (Round 1 cut the
        # spectrum into ky slabs and was slower than the replicated solve on two GPUs; the kx-slab exchange of round 2
        # moves half the bytes and runs the z pass locally: 4.41 ms against 4.64 ms replicated on two GPUs, config 5.)
        self.sharded = False
        if group is not None and sharded_solve is None:
            import torch.distributed as dist
            sharded_solve = dist.get_world_size(group) >= 2
        if group is not None and sharded_solve:
            import torch.distributed as dist
            w = dist.get_world_size(group)
            ly = 8
            while ly < 2 * grid_size[1]:
                ly *= 2
            self.sharded = w > 1 and grid_size[2] % w == 0 and ly % w == 0 and dist.get_backend(group) == "nccl"
        nx, ny, nz = grid_size
        td = _torch_dtype(npdt)
        dev = "cuda:%d" % self.device
        self._rho = torch.zeros((nz, ny, nx), dtype=td, device=dev)
        self._efield = torch.zeros((3, nz, ny, nx), dtype=td, device=dev)
        self._phi = None        # extension: allocated by solve_potential_
        self._workspace = None

    @staticmethod
    def _auto_bounds(grid_size, px, py, pz, Tn, device, handle, group=None):
        """src/mesh.jl:118-156: extrema and first delta in the particles' precision, the 1e-6
        padding and the final delta in Float64, zero delta -> 1e-6, then the cast to T."""
        torch = _torch()
        ext = []
        if all(isinstance(p, torch.Tensor) and p.device.type == "cuda" for p in (px, py, pz)):
            hd = handle if handle is not None else default_handle(px.device.index)
            hd.use_current_stream()
            (x, sx), (y, sy), (z, sz) = (_particle_view(p, px.device.index) for p in (px, py, pz))
            if not (x.dtype == y.dtype == z.dtype):
                raise ErrorException("particle coordinate arrays must share one element type")
            omin, omax = _lib.f64x3((0, 0, 0)), _lib.f64x3((0, 0, 0))
            if sx == sy == sz == 1:
                hd.check(hd.lib.scb_bounds(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype), omin, omax))
            else:
                hd.check(hd.lib.scb_bounds_strided(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(),
                                                   C.byref(_strides(sx, sy, sz)), _tag(x.dtype), omin, omax))
            P = np.float32 if x.dtype == torch.float32 else np.float64
            ext = [(P(omin[a]), P(omax[a]), P) for a in range(3)]
        else:
            for p in (px, py, pz):
                if isinstance(p, torch.Tensor):
                    p = p.detach().cpu().numpy()
                ext.append(_extrema_host(p))
        if group is not None:
            # particle-sharded run: the bunch extrema are the extrema over all ranks
            import torch.distributed as dist
            from .sharding import global_extrema
            dev = "cuda:%d" % device if torch.cuda.is_available() and dist.get_backend(group) == "nccl" else "cpu"
            glo, ghi = global_extrema([e[0] for e in ext], [e[1] for e in ext], group, dev)
            ext = [(e[2](glo[a]), e[2](ghi[a]), e[2]) for a, e in enumerate(ext)]
        lo1, hi1, d1 = [], [], []
        for (lo, hi, P), n in zip(ext, grid_size):
            d0 = P((hi - lo) / P(n - 1))
            l = np.float64(lo) - 1e-6 * np.float64(d0)
            h = np.float64(hi) + 1e-6 * np.float64(d0)
            d = (h - l) / np.float64(n - 1)
            if d == 0:
                d = np.float64(1e-6)
            lo1.append(Tn(l))
            hi1.append(Tn(h))
            d1.append(Tn(d))
        return tuple(lo1), tuple(hi1), tuple(d1)

    # column-major views, indexable like the reference's arrays (0-based)
    @property
    def rho(self):
        return self._rho.permute(2, 1, 0)

    @property
    def efield(self):
        return self._efield.permute(3, 2, 1, 0)

    @property
    def phi(self):
        """Scalar potential (nx,ny,nz) written by ``solve_potential_`` (extension; the reference's
        Mesh3D has no such field, src/mesh.jl:19-34)."""
        if self._phi is None:
            raise ErrorException("mesh.phi is only available after solve_potential_(mesh)")
        return self._phi.permute(2, 1, 0)

    def __repr__(self):  # src/mesh.jl:240-246
        nx, ny, nz = self.grid_size
        lo, hi = self.min_bounds, self.max_bounds
        return ("Mesh3D{%s, CudaTensor} (%dx%dx%d) bounds=[(%s,%s,%s), (%s,%s,%s)] gamma=%s"
                % (np.dtype(self.T).name, nx, ny, nz, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], self.gamma))

    # helpers for the ABI calls
    def remesh_(self, particles_x, particles_y, particles_z):
        """Re-fit the bounds and spacing to new particle positions in place (what a tracking loop does
        every step by calling the particle-based constructor again, src/mesh.jl:95-174) without
        re-allocating rho / efield.  The extrema come from one fused device reduction (scb_bounds); the
        next solve_ rebuilds the Green spectrum for the new spacing (cold-geometry path)."""
        npdt = np.dtype(self.T).type
        old = (self.min_bounds, self.max_bounds, self.delta)
        self.min_bounds, self.max_bounds, self.delta = self._auto_bounds(
            self.grid_size, particles_x, particles_y, particles_z, npdt, self.device, self.handle, self.group)
        if (self.min_bounds, self.max_bounds, self.delta) != old:
            # a tracking loop never returns to an old spacing: hand the old spectrum's buffer back to the
            # handle's pool so that the rebuild reuses it (no cudaMalloc / cudaFree in steady state)
            self.handle.drop_green_cache()
        return self

    def reduce_rho_(self):
        """Sum the per-rank partial charge grids in place (sharded mode keeps them partial)."""
        if self.group is not None:
            _allreduce_rho(self)

    def _n(self):
        return _lib.i64x3(self.grid_size)

    def _lo(self):
        return _lib.f64x3(self.min_bounds)

    def _hi(self):
        return _lib.f64x3(self.max_bounds)

    def _d(self):
        return _lib.f64x3(self.delta)

    def _mdt(self):
        return _lib.SCB_F32 if self.T == np.float32 else _lib.SCB_F64


# -------------------------------------------------------------------------------- operations
def _allreduce_rho(mesh: Mesh3D) -> None:
    """Sum of the ranks' charge grids: the library's own NCCL communicator on the handle's stream
    (scb_allreduce_rho) for NCCL groups; torch.distributed only for the gloo groups of the CPU-side tests."""
    import torch.distributed as dist
    if dist.get_backend(mesh.group) == "nccl":
        hd = mesh.handle
        hd.use_current_stream()
        hd.init_comm(mesh.group)
        hd.check(hd.lib.scb_allreduce_rho(hd.h, mesh._rho.data_ptr(), mesh._n(), mesh._mdt()))
    else:
        from .sharding import allreduce_rho
        allreduce_rho(mesh._rho, mesh.group)


def clear_mesh_(mesh: Mesh3D) -> None:
    """clear_mesh!  (src/deposition.jl:10-12)"""
    hd = mesh.handle
    hd.use_current_stream()
    hd.check(hd.lib.scb_clear(hd.h, mesh._rho.data_ptr(), mesh._n(), mesh._mdt()))


def deposit_(mesh: Mesh3D, particles_x, particles_y, particles_z, particles_q, clear: bool = True) -> None:
    """deposit!  (src/deposition.jl:218-247).  With ``mesh.group`` set (particle-sharded run) the
    local charge grids are summed over the ranks afterwards."""
    if not (len(particles_x) == len(particles_y) == len(particles_z) == len(particles_q)):
        raise ErrorException("Particle coordinate and charge arrays must have the same length.")
    hd = mesh.handle
    hd.use_current_stream()
    (x, sx), (y, sy), (z, sz) = (_particle_view(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    q, sq = _particle_view(particles_q, mesh.device, allow_broadcast=True)
    if not (x.dtype == y.dtype == z.dtype == q.dtype):
        raise ErrorException("particle arrays must share one element type")
    if mesh.group is not None and not mesh.sharded and not clear:
        # rho already holds a grid summed over the ranks: only THIS call's contribution may be all-reduced
        # (reducing the whole grid again would count the earlier charge once per rank)
        keep = mesh._rho.clone()
        deposit_(mesh, particles_x, particles_y, particles_z, particles_q, clear=True)   # reduces the new part
        mesh._rho.add_(keep)
        return
    if sx == sy == sz == sq == 1:
        hd.check(hd.lib.scb_deposit(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(), _tag(x.dtype),
                                    mesh._rho.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(), 1 if clear else 0))
    else:   # strided views (AoS records, broadcast charge): read in place, no dense copy
        hd.check(hd.lib.scb_deposit_strided(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(),
                                            C.byref(_strides(sx, sy, sz, sq)), _tag(x.dtype), mesh._rho.data_ptr(),
                                            mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(), 1 if clear else 0))
    if mesh.group is not None and not mesh.sharded:
        _allreduce_rho(mesh)


def solve_(mesh: Mesh3D, at_cathode: bool = False) -> None:
    """solve!  (src/solvers/free_space.jl:14-47)"""
    hd = mesh.handle
    hd.use_current_stream()
    if mesh.sharded:
        hd.init_comm(mesh.group)
        hd.check(hd.lib.scb_solve_sharded(hd.h, mesh._rho.data_ptr(), mesh._efield.data_ptr(), mesh._mdt(), mesh._n(),
                                          mesh._lo(), mesh._hi(), mesh._d(), float(mesh.gamma), 1 if at_cathode else 0))
        return
    hd.check(hd.lib.scb_solve(hd.h, mesh._rho.data_ptr(), mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(),
                              mesh._hi(), mesh._d(), float(mesh.gamma), 1 if at_cathode else 0))


def solve_potential_(mesh: Mesh3D, at_cathode: bool = False) -> None:
    """Extension: ``solve!`` that also fills ``mesh.phi`` with the scalar potential,
    computed with the reference's ``potential_green_function`` (src/green_functions.jl:13-22) as a fourth
    component of the same fused convolution.  ``mesh.efield`` is written exactly as by ``solve_``."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    if mesh._phi is None:
        mesh._phi = torch.zeros_like(mesh._rho)
    rho = mesh._rho
    if mesh.sharded:
        # the slab-decomposed solve carries E only; for the potential the partial grids are summed into a scratch copy
        # and every rank runs the four-component solve on the full grid (mesh.rho keeps this rank's partial grid)
        import torch.distributed as dist
        rho = mesh._rho.clone()
        if dist.get_backend(mesh.group) == "nccl":
            hd.init_comm(mesh.group)
            hd.check(hd.lib.scb_allreduce_rho(hd.h, rho.data_ptr(), mesh._n(), mesh._mdt()))
        else:
            dist.all_reduce(rho, op=dist.ReduceOp.SUM, group=mesh.group)
    hd.check(hd.lib.scb_solve_potential(hd.h, rho.data_ptr(), mesh._efield.data_ptr(), mesh._phi.data_ptr(),
                                        mesh._mdt(), mesh._n(), mesh._lo(), mesh._hi(), mesh._d(), float(mesh.gamma),
                                        1 if at_cathode else 0))


def magnetic_field(mesh: Mesh3D):
    """Extension: B = (beta/c) z_hat x E of a bunch moving along +z with ``mesh.gamma``; returns a
    column-major (nx,ny,nz,3) view like ``mesh.efield``."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    b = torch.empty_like(mesh._efield)
    hd.check(hd.lib.scb_bfield(hd.h, mesh._efield.data_ptr(), b.data_ptr(), mesh._mdt(), mesh._n(), float(mesh.gamma)))
    return b.permute(3, 2, 1, 0)


def interpolate_kick_(mesh: Mesh3D, particles_x, particles_y, particles_z, px, py, pz, coef_xy: float, coef_z: float) -> None:
    """Extension: ``interpolate_field`` fused with the momentum update
    ``p += coef * E`` -- the interpolated field is never written to memory.  ``px, py, pz`` are CUDA
    tensors of the particles' element type, updated in place."""
    hd = mesh.handle
    hd.use_current_stream()
    (x, sx), (y, sy), (z, sz) = (_particle_view(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    if not (x.dtype == y.dtype == z.dtype == px.dtype == py.dtype == pz.dtype):
        raise ErrorException("particle arrays must share one element type")
    if not (x.numel() == px.numel() == py.numel() == pz.numel()):
        raise ErrorException("Particle coordinate and momentum arrays must have the same length.")
    for p in (px, py, pz):
        if p.device.type != "cuda" or p.dim() != 1 or (p.numel() > 1 and p.stride(0) < 1):
            raise ErrorException("momentum arrays must be 1-D CUDA tensors or positive-stride views (they are updated in place)")
    so = [p.stride(0) if p.numel() > 1 else 1 for p in (px, py, pz)]
    if sx == sy == sz == 1 and so == [1, 1, 1]:
        hd.check(hd.lib.scb_interpolate_kick(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype),
                                             mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(),
                                             px.data_ptr(), py.data_ptr(), pz.data_ptr(), float(coef_xy), float(coef_z)))
    else:
        hd.check(hd.lib.scb_interpolate_kick_strided(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(),
                                                     C.byref(_strides(sx, sy, sz, 1, *so)), _tag(x.dtype),
                                                     mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(),
                                                     px.data_ptr(), py.data_ptr(), pz.data_ptr(), float(coef_xy), float(coef_z)))


def solve_freespace_(mesh: Mesh3D, offset=(0.0, 0.0, 0.0)) -> None:
    """solve_freespace!  (src/solvers/free_space.jl:56-101)"""
    hd = mesh.handle
    hd.use_current_stream()
    Tn = np.dtype(mesh.T).type
    off = tuple(float(Tn(v)) for v in offset)
    hd.check(hd.lib.scb_solve_freespace(hd.h, mesh._rho.data_ptr(), mesh._efield.data_ptr(), mesh._mdt(), mesh._n(),
                                        mesh._d(), float(mesh.gamma), _lib.f64x3(off)))


def interpolate_field(mesh: Mesh3D, particles_x, particles_y, particles_z):
    """interpolate_field  (src/interpolation.jl:100-128): returns (Ex, Ey, Ez) with the particles'
    element type, freshly allocated like ``similar(particles_x)``."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    (x, sx), (y, sy), (z, sz) = (_particle_view(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    if not (x.dtype == y.dtype == z.dtype):
        raise ErrorException("particle arrays must share one element type")
    ex, ey, ez = (torch.empty(x.numel(), dtype=x.dtype, device=x.device) for _ in range(3))
    if sx == sy == sz == 1:
        hd.check(hd.lib.scb_interpolate(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype),
                                        mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(),
                                        ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))
    else:
        hd.check(hd.lib.scb_interpolate_strided(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(),
                                                C.byref(_strides(sx, sy, sz)), _tag(x.dtype), mesh._efield.data_ptr(),
                                                mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(),
                                                ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))
    return ex, ey, ez


def get_green_function_(shape2, delta, gamma, icomp: int, offset=(0.0, 0.0, 0.0), T=np.float64,
                        handle: Optional[Handle] = None):
    """get_green_function!  (src/green_functions.jl:41-67): the real part of the reference's cgrn
    array of shape (2nx, 2ny, 2nz), as a column-major torch view."""
    torch = _torch()
    hd = handle if handle is not None else default_handle()
    hd.use_current_stream()
    npdt = _np_dtype(T)
    sx, sy, sz = (int(v) for v in shape2)
    out = torch.empty((sz, sy, sx), dtype=_torch_dtype(npdt), device="cuda:%d" % hd.device)
    hd.check(hd.lib.scb_green(hd.h, out.data_ptr(), _lib.i64x3((sx, sy, sz)), _lib.f64x3(delta), float(gamma), int(icomp),
                              _lib.f64x3(offset), _lib.SCB_F32 if npdt == np.float32 else _lib.SCB_F64))
    return out.permute(2, 1, 0)


def cell_indices(mesh: Mesh3D, particles_x, particles_y, particles_z):
    """Parity hook: floor((p - min) / delta) per axis as int64."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    x, y, z = (_device_array(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    ix, iy, iz = (torch.empty(x.numel(), dtype=torch.int64, device=x.device) for _ in range(3))
    hd.check(hd.lib.scb_cell_index(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype), mesh._mdt(),
                                   mesh._lo(), mesh._d(), ix.data_ptr(), iy.data_ptr(), iz.data_ptr()))
    return ix, iy, iz


# ------------------------------------------------------------------ bunches kept ordered by cell
def set_particle_order(mesh_or_handle, order) -> None:
    """Tell the handle how the caller's bunch is ordered: ``"random"`` (default) or ``"cell"`` (the bunch was put in
    cell order by ``sort_particles_`` and is re-sorted every few steps).  Results do not depend on the setting, only
    the speed of deposit_ / interpolate_field / step_ does."""
    hd = mesh_or_handle.handle if isinstance(mesh_or_handle, Mesh3D) else mesh_or_handle
    hd.set_particle_order(order)


def sort_particles(mesh: Mesh3D, particles_x, particles_y, particles_z, out=None):
    """Permutation that orders the bunch by linear cell index of ``mesh`` (scb_sort_particles; stable).  Returns an
    int32 CUDA tensor ``perm``: ``x[perm.long()]`` is the ordered array (``permute`` does that in one pass).
    ``out``: an int32 CUDA tensor to receive the permutation (a tracking loop reuses it across re-sorts)."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    x, y, z = (_device_array(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    if not (x.dtype == y.dtype == z.dtype):
        raise ErrorException("particle arrays must share one element type")
    if not (x.numel() == y.numel() == z.numel()):
        raise ErrorException("Particle coordinate arrays must have the same length.")
    perm = out if out is not None else torch.empty(x.numel(), dtype=torch.int32, device=x.device)
    if perm.dtype != torch.int32 or perm.numel() != x.numel() or perm.device != x.device or not perm.is_contiguous():
        raise ErrorException("sort_particles: out must be a contiguous int32 CUDA tensor of the particles' length")
    hd.check(hd.lib.scb_sort_particles(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype),
                                       mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(), perm.data_ptr()))
    return perm


def permute(perm, *arrays, handle: Optional[Handle] = None, out=None):
    """``tuple(a[perm] for a in arrays)`` in passes of up to 8 arrays (scb_permute): 1-D CUDA tensors of one
    floating-point type and the length of ``perm``.  ``out``: destination tensors (same types and lengths, not
    aliasing the sources) instead of fresh allocations."""
    torch = _torch()
    if not arrays:
        return ()
    hd = handle if handle is not None else default_handle(perm.device.index)
    hd.use_current_stream()
    dt = arrays[0].dtype
    n = perm.numel()
    for a in arrays:
        if a.dtype != dt or a.numel() != n or a.device != perm.device or not a.is_contiguous():
            raise ErrorException("permute: arrays must be contiguous CUDA tensors of one type and the length of perm")
    outs = list(out) if out is not None else [torch.empty_like(a) for a in arrays]
    if len(outs) != len(arrays) or any(o.dtype != dt or o.numel() != n or o.device != perm.device or not o.is_contiguous()
                                       for o in outs):
        raise ErrorException("permute: out must match the arrays in number, type, length and device")
    for first in range(0, len(arrays), 8):
        src = arrays[first:first + 8]
        dst = outs[first:first + 8]
        sp = (C.c_void_p * len(src))(*[a.data_ptr() for a in src])
        dp = (C.c_void_p * len(dst))(*[a.data_ptr() for a in dst])
        hd.check(hd.lib.scb_permute(hd.h, n, perm.data_ptr(), len(src), sp, dp, _tag(dt)))
    return tuple(outs)


def sort_particles_(mesh: Mesh3D, particles_x, particles_y, particles_z, *others):
    """Order a bunch by cell: returns ``(perm, x, y, z, *others)`` with every array permuted (new tensors)."""
    x, y, z = (_device_array(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    perm = sort_particles(mesh, x, y, z)
    return (perm,) + permute(perm, x, y, z, *[_device_array(a, mesh.device) for a in others], handle=mesh.handle)


def particle_order_fraction(mesh: Mesh3D, particles_x, particles_y, particles_z) -> float:
    """Fraction of sampled neighbouring particle pairs that share a cell or sit in x-adjacent cells
    (scb_particle_order_fraction): about 1 for a cell-ordered bunch, about 0 for a random one."""
    hd = mesh.handle
    hd.use_current_stream()
    x, y, z = (_device_array(a, mesh.device) for a in (particles_x, particles_y, particles_z))
    out = C.c_double(0.0)
    hd.check(hd.lib.scb_particle_order_fraction(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype),
                                                mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(), C.byref(out)))
    return float(out.value)


def step_(mesh: Mesh3D, x, y, z, q, ex, ey, ez, at_cathode: bool = False) -> None:
    """deposit! + solve! + interpolate_field on device-resident tensors with caller-owned outputs
    (the timed body of benchmark/full_pipeline_benchmark.jl:26-30, without the per-call allocation)."""
    torch = _torch()
    hd = mesh.handle
    hd.use_current_stream()
    for t in (x, y, z, q, ex, ey, ez):
        if not isinstance(t, torch.Tensor) or t.device.type != "cuda" or t.dim() != 1:
            raise ErrorException("step_ takes 1-D CUDA tensors")
        if t.dtype != x.dtype:
            raise ErrorException("particle arrays must share one element type")
        if t.numel() != x.numel():
            raise ErrorException("Particle coordinate and charge arrays must have the same length.")
        if t.device != x.device:
            raise ErrorException("particle arrays must live on one device")
    st = [t.stride(0) if t.numel() > 1 else 1 for t in (x, y, z, q, ex, ey, ez)]
    if st != [1] * 7:
        if mesh.group is not None:
            raise ErrorException("strided particle views are single-GPU; pass dense shards to a particle-sharded step_")
        if min(st[:3] + st[4:]) < 1 or st[3] < 0:
            raise ErrorException("particle views need positive strides (the charge may be a stride-0 broadcast)")
        hd.check(hd.lib.scb_step_strided(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(),
                                         C.byref(_strides(*st)), _tag(x.dtype), mesh._rho.data_ptr(),
                                         mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._hi(), mesh._d(),
                                         float(mesh.gamma), 1 if at_cathode else 0, ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))
        return
    if mesh.group is None:
        hd.check(hd.lib.scb_step(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(), _tag(x.dtype),
                                 mesh._rho.data_ptr(), mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(),
                                 mesh._hi(), mesh._d(), float(mesh.gamma), 1 if at_cathode else 0,
                                 ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))
        return
    if mesh.sharded:
        # slab-decomposed solve with the all-gather of the field overlapped with the interpolation
        hd.init_comm(mesh.group)
        hd.check(hd.lib.scb_step_sharded(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(),
                                         _tag(x.dtype), mesh._rho.data_ptr(), mesh._efield.data_ptr(), mesh._mdt(),
                                         mesh._n(), mesh._lo(), mesh._hi(), mesh._d(), float(mesh.gamma),
                                         1 if at_cathode else 0, ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))
        return
    deposit_(mesh, x, y, z, q)
    solve_(mesh, at_cathode=at_cathode)
    hd.check(hd.lib.scb_interpolate(hd.h, x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), _tag(x.dtype),
                                    mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._d(),
                                    ex.data_ptr(), ey.data_ptr(), ez.data_ptr()))


def _host_ptr(a):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        if a.device.type != "cpu" or not a.is_contiguous() or a.dtype not in (torch.float32, torch.float64):
            raise ErrorException("host-buffer steps take contiguous Float32/Float64 CPU arrays")
        return a.data_ptr(), a.dtype
    if not isinstance(a, np.ndarray) or not a.flags["C_CONTIGUOUS"] or a.dtype not in (np.float32, np.float64):
        raise ErrorException("host-buffer steps take contiguous Float32/Float64 CPU arrays")
    return a.ctypes.data, (torch.float32 if a.dtype == np.float32 else torch.float64)


def _step_host(fn_name, mesh, x, y, z, q, ex, ey, ez, at_cathode):
    hd = mesh.handle
    hd.use_current_stream()
    ptrs = [_host_ptr(a) for a in (x, y, z, q, ex, ey, ez)]
    if any(d != ptrs[0][1] for _, d in ptrs):
        raise ErrorException("particle arrays must share one element type")
    if any(len(a) != len(x) for a in (y, z, q, ex, ey, ez)):
        raise ErrorException("Particle coordinate and charge arrays must have the same length.")
    (px, dt), (py, _), (pz, _), (pq, _), (pex, _), (pey, _), (pez, _) = ptrs
    if mesh.group is not None:
        # particle shards: every rank feeds its own host shard (scb_step_host_sharded_async)
        hd.init_comm(mesh.group)
        hd.check(hd.lib.scb_step_host_sharded_async(hd.h, len(x), px, py, pz, pq, _tag(dt), mesh._rho.data_ptr(),
                                                    mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._hi(),
                                                    mesh._d(), float(mesh.gamma), 1 if at_cathode else 0,
                                                    1 if mesh.sharded else 0, pex, pey, pez))
        if fn_name == "scb_step_host":
            hd.check(hd.lib.scb_step_host_wait(hd.h))
        return
    hd.check(getattr(hd.lib, fn_name)(hd.h, len(x), px, py, pz, pq, _tag(dt), mesh._rho.data_ptr(),
                                      mesh._efield.data_ptr(), mesh._mdt(), mesh._n(), mesh._lo(), mesh._hi(), mesh._d(),
                                      float(mesh.gamma), 1 if at_cathode else 0, pex, pey, pez))


def step_host_(mesh: Mesh3D, x, y, z, q, ex, ey, ez, at_cathode: bool = False) -> None:
    """The same step with HOST particle buffers (numpy arrays or CPU torch tensors, ideally
    pinned): host->device and device->host copies are part of the call (scb_step_host).  On a mesh built
    with ``group=`` the buffers are this rank's particle shard (scb_step_host_sharded_async + wait)."""
    _step_host("scb_step_host", mesh, x, y, z, q, ex, ey, ez, at_cathode)


def step_host_async_(mesh: Mesh3D, x, y, z, q, ex, ey, ez, at_cathode: bool = False) -> None:
    """Queue a host-buffer step and return (scb_step_host_async).  Up to two steps may be in flight: the
    upload of one overlaps the download of the previous one.  The output buffers of consecutive steps
    must be distinct and every buffer must stay alive and untouched until step_host_wait_."""
    _step_host("scb_step_host_async", mesh, x, y, z, q, ex, ey, ez, at_cathode)


def step_host_wait_(mesh: Mesh3D) -> None:
    """Wait for every queued host-buffer step (scb_step_host_wait)."""
    hd = mesh.handle
    hd.check(hd.lib.scb_step_host_wait(hd.h))
