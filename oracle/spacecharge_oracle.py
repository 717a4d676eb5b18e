"""CPU oracle for the space-charge hot path (deposit -> IGF/FFT solve -> interpolate).

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of the algorithm
of bmad-sim/SpaceCharge.jl v1.2.0 for the path BASELINE.json names.  It may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
(``spacecharge.jl_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference ships no golden vectors or known-answer arrays
for this path and Julia is not installed in this image,
so the reference itself cannot be executed to produce fixtures.  The oracle is
pinned instead against (i) every behavioural assertion the reference's own
tests make (tests/test_oracle_reference_asserts.py), (ii) a direct O(Ng^2)
summation of rho * IGF with the IGF evaluated in 50-digit arithmetic
(tests/test_oracle_direct_sum.py) and (iii) the analytic isotropic-Gaussian
field (reference test/analytical_test.jl).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Arrays use the reference's layout: ``rho[ix, iy, iz]`` and
``efield[ix, iy, iz, c]`` stored column-major (x fastest, component slowest),
i.e. NumPy arrays created with ``order='F'``.

Third-party arithmetic that is not under the reference tree:
  * FFT: AbstractFFTs 1.5.0 / FFTW.jl 1.9.0 / FFTW_jll 3.3.11 (Manifest.toml);
    semantics used = unnormalised forward DFT, inverse scaled by 1/M.  Restated
    with scipy.fft (pocketfft), any correct FFT agrees to round-off.
  * libm (atan/log/sqrt) from Julia Base; restated with NumPy's libm calls.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import scipy.fft as _sfft

# src/utils.jl:7-8
CLIGHT = 299792458.0
FPEI = CLIGHT ** 2 * 1.0e-7  # 1/(4 pi eps0) = 8.987551787368176e9


def _workers() -> int:
    return int(os.environ.get("SCB_ORACLE_THREADS", os.cpu_count() or 1))


class ErrorException(Exception):
    """Mirror of Julia's ErrorException raised by ``error("...")``."""


# --------------------------------------------------------------------------- mesh
@dataclass
class Mesh3D:
    """src/mesh.jl:19-34 (struct fields, same names)."""

    grid_size: Tuple[int, int, int]
    min_bounds: Tuple[float, float, float]
    max_bounds: Tuple[float, float, float]
    delta: Tuple[float, float, float]
    gamma: float
    total_charge: float
    rho: np.ndarray
    efield: np.ndarray
    T: type = np.float64
    _workspace: Optional[dict] = field(default=None, repr=False)
    phi: Optional[np.ndarray] = field(default=None, repr=False)   # extension, see solve(..., potential=True)


def _alloc(grid_size, T):
    nx, ny, nz = grid_size
    rho = np.zeros((nx, ny, nz), dtype=T, order="F")
    efield = np.zeros((nx, ny, nz, 3), dtype=T, order="F")
    return rho, efield


def auto_bounds(grid_size, px, py, pz, T=np.float64):
    """Bounds/delta arithmetic of ctor #1, src/mesh.jl:118-156.

    extrema and the first delta are in the particles' precision P (:120-129);
    the 1e-6 padding multiplies by a Float64 literal, so everything after it is
    Float64 (:132-142); zero delta -> 1e-6 (:145-149); only then cast to T
    (:150-156).  Returns (min_bounds, max_bounds, delta) as tuples of T scalars.
    """
    lo0, hi0, d0 = [], [], []
    for p, n in zip((px, py, pz), grid_size):
        p = np.asarray(p)
        P = p.dtype.type if p.dtype.kind == "f" else np.float64
        lo = P(p.min())
        hi = P(p.max())
        lo0.append(lo)
        hi0.append(hi)
        d0.append(P((hi - lo) / P(n - 1)))
    lo1 = [np.float64(l) - 1e-6 * np.float64(d) for l, d in zip(lo0, d0)]
    hi1 = [np.float64(h) + 1e-6 * np.float64(d) for h, d in zip(hi0, d0)]
    d1 = [(h - l) / np.float64(n - 1) for h, l, n in zip(hi1, lo1, grid_size)]
    d1 = [np.float64(1e-6) if d == 0 else d for d in d1]
    T = np.dtype(T).type
    return (tuple(T(v) for v in lo1), tuple(T(v) for v in hi1), tuple(T(v) for v in d1))


def mesh_from_particles(grid_size, px, py, pz, T=np.float64, gamma=1.0, total_charge=0.0) -> Mesh3D:
    """ctor #1, src/mesh.jl:95-174."""
    grid_size = tuple(int(g) for g in grid_size)
    if any(g <= 1 for g in grid_size):  # :106-108
        raise ErrorException("All elements of grid_size must be at least 2.")
    if len(px) == 0 or len(py) == 0 or len(pz) == 0:  # :110-112
        raise ErrorException("Particle arrays cannot be empty.")
    if not (len(px) == len(py) == len(pz)):  # :114-116
        raise ErrorException("Particle coordinate arrays must have the same length.")
    T = np.dtype(T).type
    lo, hi, d = auto_bounds(grid_size, px, py, pz, T)
    rho, efield = _alloc(grid_size, T)
    return Mesh3D(grid_size, lo, hi, d, T(gamma), T(total_charge), rho, efield, T)


def mesh_from_bounds(grid_size, min_bounds, max_bounds, T=np.float64, gamma=1.0, total_charge=0.0) -> Mesh3D:
    """ctor #2, src/mesh.jl:196-238: cast bounds to T first, delta computed in T (:214-220)."""
    grid_size = tuple(int(g) for g in grid_size)
    if any(g <= 1 for g in grid_size):  # :206-208
        raise ErrorException("All elements of grid_size must be at least 2.")
    if any(h <= l for h, l in zip(max_bounds, min_bounds)):  # :209-211
        raise ErrorException("max_bounds must be strictly greater than min_bounds for all dimensions.")
    T = np.dtype(T).type
    lo = tuple(T(v) for v in min_bounds)
    hi = tuple(T(v) for v in max_bounds)
    d = tuple(T((h - l) / T(n - 1)) for h, l, n in zip(hi, lo, grid_size))
    rho, efield = _alloc(grid_size, T)
    return Mesh3D(grid_size, lo, hi, d, T(gamma), T(total_charge), rho, efield, T)


# ------------------------------------------------------------------ index / weights
def cell_index_and_frac(p, lo, delta, clamp_n=None):
    """src/deposition.jl:39-51 == src/interpolation.jl:31-43.

    t=(p-lo)/delta in promote(P,T) with a true division; i=floor(Int,t); d=t-i.
    ``clamp_n`` (not in the reference): clamp i to [0, clamp_n-2] *before* forming d, which is what
    the CUDA kernels do -- identical for in-range particles, and a particle sitting exactly on the
    upper bound lands on node n-1 with weight 1 instead of writing out of bounds.
    """
    p = np.asarray(p)
    W = np.promote_types(p.dtype, np.asarray(lo).dtype)
    t = (p.astype(W) - W.type(lo)) / W.type(delta)
    i = np.floor(t).astype(np.int64)
    if clamp_n is not None:
        i = np.clip(i, 0, clamp_n - 2)
    d = t - i.astype(W)
    return i, d


def clear_mesh(mesh: Mesh3D) -> None:
    """src/deposition.jl:10-12."""
    mesh.rho.fill(0)


def deposit(mesh: Mesh3D, px, py, pz, pq, clear: bool = True, clamp: bool = False) -> None:
    """deposit! src/deposition.jl:218-247 with the serial CPU driver (:167-197).

    Corner value ((q*wx)*wy)*wz, eight updates in the order of :77-84, particles
    accumulated serially into rho of type T.  ``np.add.at`` is unbuffered and
    applies the updates in index order, which reproduces the serial loop when the
    update list is laid out particle-major / corner-minor as done here.
    ``clamp=True`` applies the [0, n-2] index clamp of the CUDA path (identical
    for in-range particles).
    """
    if not (len(px) == len(py) == len(pz) == len(pq)):  # :226-228
        raise ErrorException("Particle coordinate and charge arrays must have the same length.")
    if clear:
        clear_mesh(mesh)
    nx, ny, nz = mesh.grid_size
    ix, dx = cell_index_and_frac(px, mesh.min_bounds[0], mesh.delta[0], nx if clamp else None)
    iy, dy = cell_index_and_frac(py, mesh.min_bounds[1], mesh.delta[1], ny if clamp else None)
    iz, dz = cell_index_and_frac(pz, mesh.min_bounds[2], mesh.delta[2], nz if clamp else None)
    one = dx.dtype.type(1)
    wx = (one - dx, dx)
    wy = (one - dy, dy)
    wz = (one - dz, dz)
    q = np.asarray(pq)
    W = np.promote_types(q.dtype, dx.dtype)
    q = q.astype(W)
    npart = len(q)
    lin = np.empty((npart, 8), dtype=np.int64)
    val = np.empty((npart, 8), dtype=W)
    k = 0
    for c in (0, 1):  # z corner slowest, x fastest: order of src/deposition.jl:77-84
        for b in (0, 1):
            for a in (0, 1):
                lin[:, k] = (ix + a) + nx * ((iy + b) + ny * (iz + c))
                val[:, k] = q * wx[a] * wy[b] * wz[c]
                k += 1
    flat = mesh.rho.reshape(-1, order="F")
    if flat.dtype == W:
        np.add.at(flat, lin.reshape(-1), val.reshape(-1))
    else:
        # T narrower than the addend: rho + v is formed in promote(T, W) and rounded
        # to T at every store, exactly like the Julia loop.
        li = lin.reshape(-1)
        va = val.reshape(-1)
        for j in range(li.size):
            flat[li[j]] = flat.dtype.type(W.type(flat[li[j]]) + va[j])
    mesh.rho[...] = flat.reshape(mesh.rho.shape, order="F")


# ------------------------------------------------------------------ Green functions
def potential_green_function(x, y, z):
    """src/green_functions.jl:13-22 (defined by the reference, unreachable from solve!)."""
    r = np.sqrt(x * x + y * y + z * z)
    half = x.dtype.type(0.5) if hasattr(x, "dtype") else 0.5
    with np.errstate(all="ignore"):
        v = (-half * z * z * np.arctan(x * y / (z * r)) - half * y * y * np.arctan(x * z / (y * r))
             - half * x * x * np.arctan(y * z / (x * r)) + y * z * np.log(x + r)
             + x * z * np.log(y + r) + x * y * np.log(z + r))
    return np.where(r == 0, 0, v)


def field_green_function(x, y, z):
    """src/green_functions.jl:35-38."""
    r = np.sqrt(x * x + y * y + z * z)
    return x * np.arctan((y * z) / (r * x)) - z * np.log(r + y) + y * np.log((r - z) / (r + z)) / 2


def green_pointwise(shape2, delta, gamma, icomp, offset=(0.0, 0.0, 0.0), T=np.float64, potential_icomp0=False):
    """get_green_kernel!, src/green_functions.jl:69-101 (real part; imaginary part is zero)."""
    T = np.dtype(T).type
    isize, jsize, ksize = shape2
    gamma = T(gamma)
    dx, dy, dz = T(delta[0]), T(delta[1]), T(T(delta[2]) * gamma)
    if icomp in (1, 2):
        factor = T(gamma / T(T(dx * dy) * dz))
    else:
        factor = T(T(1) / T(T(dx * dy) * dz))
    umin = T(T(T(1 - isize) * dx) / T(2)) + T(offset[0])
    vmin = T(T(T(1 - jsize) * dy) / T(2)) + T(offset[1])
    wmin = T(T(T(1 - ksize) * dz) / T(2)) + T(T(offset[2]) * gamma)
    u = (np.arange(isize).astype(T) * dx + umin).astype(T)[:, None, None]
    v = (np.arange(jsize).astype(T) * dy + vmin).astype(T)[None, :, None]
    w = (np.arange(ksize).astype(T) * dz + wmin).astype(T)[None, None, :]
    u, v, w = np.broadcast_arrays(u, v, w)
    with np.errstate(all="ignore"):
        if icomp == 1:
            g = field_green_function(u, v, w) * factor
        elif icomp == 2:
            g = field_green_function(v, w, u) * factor
        elif icomp == 3:
            g = field_green_function(w, u, v) * factor
        elif icomp == 0 and potential_icomp0:
            # EXTENSION (not reachable in the reference, whose kernel returns zero here): the potential
            # Green function with the factor 1/(dx*dy*dz) of the `else` branch of :77-81
            g = potential_green_function(u, v, w) * factor
        else:
            g = np.zeros((isize, jsize, ksize), dtype=T)
    return np.asfortranarray(g.astype(T))


def difference_8point(c):
    """apply_8point_differencing!, src/green_functions.jl:103-112 (left-to-right order)."""
    return (c[1:, 1:, 1:] - c[:-1, 1:, 1:] - c[1:, :-1, 1:] - c[1:, 1:, :-1]
            - c[:-1, :-1, :-1] + c[:-1, :-1, 1:] + c[:-1, 1:, :-1] + c[1:, :-1, :-1])


def get_green_function(shape2, delta, gamma, icomp, offset=(0.0, 0.0, 0.0), T=np.float64, potential_icomp0=False):
    """get_green_function!, src/green_functions.jl:41-67.

    Point-wise fill, 8-point differencing into the leading (2n-1)^3 block; the last
    plane of every dimension keeps its raw point-wise values (:64-66).
    """
    g = green_pointwise(shape2, delta, gamma, icomp, offset, T, potential_icomp0)
    out = g.copy(order="F")
    out[:-1, :-1, :-1] = difference_8point(g)
    return out


# --------------------------------------------------------------------------- solver
def _fftn(a, inverse=False):
    f = _sfft.ifftn if inverse else _sfft.fftn
    return f(a, workers=_workers())


def solve_freespace(mesh: Mesh3D, offset=(0.0, 0.0, 0.0), potential: bool = False) -> None:
    """solve_freespace!, src/solvers/free_space.jl:56-101 (same structure: padded C2C
    FFT of rho, then per component IGF -> FFT -> multiply -> inverse FFT -> extract).

    ``potential=True`` (EXTENSION, no reference counterpart -- parity unpinned): the same loop body run
    once more with icomp = 0 -> potential_green_function, result stored in ``mesh.phi``."""
    T = mesh.T
    CT = np.complex64 if T == np.float32 else np.complex128
    nx, ny, nz = mesh.grid_size
    crho = np.zeros((2 * nx, 2 * ny, 2 * nz), dtype=CT, order="F")  # :68
    crho[:nx, :ny, :nz] = mesh.rho  # :69
    crho = _fftn(crho).astype(CT, copy=False)  # :72
    factr = T(FPEI)  # :75
    for icomp in (1, 2, 3):  # :77
        cgrn = get_green_function((2 * nx, 2 * ny, 2 * nz), mesh.delta, mesh.gamma, icomp, offset, T)
        cgrn = _fftn(cgrn.astype(CT)).astype(CT, copy=False)  # :89
        temp = (crho * cgrn).astype(CT, copy=False)  # :92
        temp = _fftn(temp, inverse=True).astype(CT, copy=False)  # :95
        mesh.efield[:, :, :, icomp - 1] = factr * temp.real[nx - 1:2 * nx - 1, ny - 1:2 * ny - 1,
                                                            nz - 1:2 * nz - 1].astype(T)  # :98-99
    if potential:
        cgrn = get_green_function((2 * nx, 2 * ny, 2 * nz), mesh.delta, mesh.gamma, 0, offset, T, potential_icomp0=True)
        cgrn = _fftn(cgrn.astype(CT)).astype(CT, copy=False)
        temp = _fftn((crho * cgrn).astype(CT, copy=False), inverse=True).astype(CT, copy=False)
        mesh.phi = np.asfortranarray(factr * temp.real[nx - 1:2 * nx - 1, ny - 1:2 * ny - 1, nz - 1:2 * nz - 1].astype(T))


def solve(mesh: Mesh3D, at_cathode: bool = False, potential: bool = False) -> None:
    """solve!, src/solvers/free_space.jl:14-47.  ``potential``: extension, see solve_freespace."""
    T = mesh.T
    solve_freespace(mesh, (T(0), T(0), T(0)), potential)  # :17
    if at_cathode:
        rho_img, e_img = _alloc(mesh.grid_size, T)
        image = Mesh3D(mesh.grid_size, mesh.min_bounds, mesh.max_bounds, mesh.delta, mesh.gamma,
                       mesh.total_charge, rho_img, e_img, T)
        image.rho[...] = -mesh.rho[:, :, ::-1]  # :34
        offset_z = T(T(2) * mesh.min_bounds[2] + T(mesh.max_bounds[2] - mesh.min_bounds[2]))  # :39
        solve_freespace(image, (T(0), T(0), offset_z), potential)  # :42
        mesh.efield += image.efield  # :45
        if potential:
            mesh.phi += image.phi


def magnetic_field(mesh: Mesh3D) -> np.ndarray:
    """EXTENSION (no reference counterpart): B = (beta/c) z_hat x E for a bunch moving along +z."""
    T = mesh.T
    g = float(mesh.gamma)
    boc = T(np.sqrt(1.0 - 1.0 / (g * g)) / CLIGHT)
    b = np.zeros_like(mesh.efield)
    b[..., 0] = -(boc * mesh.efield[..., 1])
    b[..., 1] = boc * mesh.efield[..., 0]
    return b


# -------------------------------------------------------------------- interpolation
def interpolate_field(mesh: Mesh3D, px, py, pz, clamp: bool = False):
    """interpolate_field, src/interpolation.jl:100-128: the kernel's values stored into arrays of the
    particles' element type (:110-112)."""
    px = np.asarray(px)
    P = px.dtype if px.dtype.kind == "f" else np.float64
    return tuple(a.astype(P) for a in interpolate_field_promoted(mesh, px, py, pz, clamp))


def interpolate_field_promoted(mesh: Mesh3D, px, py, pz, clamp: bool = False):
    """interpolate_kernel!, src/interpolation.jl:17-86, values in promote(P, T) before the store.

    Weights (1-dx)*(1-dy)*(1-dz) ... left to right (:46-53); each component is the
    left-to-right sum of eight products in the order 000,100,010,110,001,101,011,111
    (:56-85); outputs have the particles' element type (:110-112).
    """
    px = np.asarray(px)
    nx, ny, nz = mesh.grid_size
    ix, dx = cell_index_and_frac(px, mesh.min_bounds[0], mesh.delta[0], nx if clamp else None)
    iy, dy = cell_index_and_frac(py, mesh.min_bounds[1], mesh.delta[1], ny if clamp else None)
    iz, dz = cell_index_and_frac(pz, mesh.min_bounds[2], mesh.delta[2], nz if clamp else None)
    one = dx.dtype.type(1)
    ax = (one - dx, dx)
    ay = (one - dy, dy)
    az = (one - dz, dz)
    out = []
    W = dx.dtype
    for c in range(3):
        e = mesh.efield[:, :, :, c]
        acc = None
        for cz in (0, 1):
            for by in (0, 1):
                for a in (0, 1):
                    w = ax[a] * ay[by] * az[cz]
                    term = e[ix + a, iy + by, iz + cz].astype(W) * w
                    acc = term if acc is None else acc + term
        out.append(acc)
    return tuple(out)


def interpolate_kick(mesh: Mesh3D, px, py, pz, mx, my, mz, coef_xy, coef_z, clamp: bool = False):
    """EXTENSION (no reference counterpart): momenta after p += coef * E(particle), product and sum
    rounded separately in promote(P, T), result cast to the particles' type."""
    e = interpolate_field_promoted(mesh, px, py, pz, clamp)
    W = e[0].dtype.type
    P = np.asarray(px).dtype
    out = []
    for m, ec, c in zip((mx, my, mz), e, (coef_xy, coef_xy, coef_z)):
        out.append((np.asarray(m).astype(W) + W(c) * ec).astype(P))
    return tuple(out)


# ------------------------------------------------------------- higher-level helpers
def full_step(grid_size, px, py, pz, pq, T=np.float64, gamma=1.0, at_cathode=False):
    """deposit! + solve! + interpolate_field on a fresh auto-bounds mesh
    (benchmark/full_pipeline_benchmark.jl:25-30)."""
    mesh = mesh_from_particles(grid_size, px, py, pz, T=T, gamma=gamma)
    deposit(mesh, px, py, pz, pq)
    solve(mesh, at_cathode=at_cathode)
    return mesh, interpolate_field(mesh, px, py, pz)


def igf_direct(dvec, delta, gamma, icomp, offset=(0.0, 0.0, 0.0)):
    """Integrated Green function for an integer displacement (field node minus source
    node), i.e. what the differenced cgrn entry at 1-based index n+d holds
    (src/green_functions.jl:82-88 with SURVEY Appendix A.4).  Float64, scalar."""
    dx, dy, dz = float(delta[0]), float(delta[1]), float(delta[2]) * float(gamma)
    fac = (float(gamma) if icomp in (1, 2) else 1.0) / (dx * dy * dz)
    cu = [(dvec[0] + s) * dx + offset[0] for s in (-0.5, 0.5)]
    cv = [(dvec[1] + s) * dy + offset[1] for s in (-0.5, 0.5)]
    cw = [(dvec[2] + s) * dz + offset[2] * float(gamma) for s in (-0.5, 0.5)]
    tot = 0.0
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                u, v, w = np.float64(cu[a]), np.float64(cv[b]), np.float64(cw[c])
                args = {1: (u, v, w), 2: (v, w, u), 3: (w, u, v)}[icomp]
                sign = (-1.0) ** (3 - a - b - c)
                tot += sign * float(field_green_function(*args)) * fac
    return tot
