"""Timed CPU baseline: the reference's algorithmic structure (C restatement in ref_port.c + a
threaded C2C FFT from scipy/pocketfft, since FFTW -- which the reference pins as FFTW_jll 3.3.11 --
is not installed).  TEST / BENCH INFRASTRUCTURE ONLY; label: "reference-algorithm restatement",
never "Julia".  Float64 only (the reference's benchmarks are Float64,
benchmark/full_pipeline_benchmark.jl:16-30)."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import scipy.fft as sfft

from . import build as _build
from .spacecharge_oracle import FPEI

_I3 = C.c_int64 * 3
_D3 = C.c_double * 3
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="A")


def _lib():
    lib = C.CDLL(_build.build())
    lib.port_deposit.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, _dp, _I3, _D3, _D3]
    lib.port_green_point.argtypes = [_dp, _I3, _D3, C.c_double, C.c_int, _D3]
    lib.port_diff8.argtypes = [_dp, _dp, _I3]
    lib.port_copy_back.argtypes = [_dp, _dp, _I3]
    lib.port_embed.argtypes = [_dp, _dp, _I3]
    lib.port_multiply.argtypes = [_dp, _dp, _dp, C.c_int64]
    lib.port_extract.argtypes = [_dp, _dp, _I3, C.c_double]
    lib.port_interpolate.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, _I3, _D3, _D3, _dp, _dp, _dp]
    for f in ("port_deposit", "port_green_point", "port_diff8", "port_copy_back", "port_embed", "port_multiply",
              "port_extract", "port_interpolate"):
        getattr(lib, f).restype = None
    lib.port_set_threads.argtypes = [C.c_int]
    lib.port_set_threads.restype = None
    lib.port_get_threads.restype = C.c_int
    return lib


class RefPort:
    """Holds the reference's workspace (three complex (2n)^3 arrays, src/mesh.jl:49-72)."""

    def __init__(self, grid, min_bounds, delta, gamma=1.0, threads=None):
        self.lib = _lib()
        self.n = tuple(int(g) for g in grid)
        self.lo = tuple(float(v) for v in min_bounds)
        self.d = tuple(float(v) for v in delta)
        self.gamma = float(gamma)
        self.threads = int(threads or os.cpu_count() or 1)
        # explicit: torchrun exports OMP_NUM_THREADS=1, which a setdefault() would leave in force (round-1 verdict)
        self.lib.port_set_threads(self.threads)
        self.omp_threads = int(self.lib.port_get_threads())
        s2 = tuple(2 * g for g in self.n)
        self.s2 = s2
        self.rho = np.zeros(self.n, order="F")
        self.efield = np.zeros(self.n + (3,), order="F")
        self.crho = np.zeros(s2, dtype=np.complex128, order="F")
        self.cgrn = np.zeros(s2, dtype=np.complex128, order="F")
        self.temp = np.zeros(s2, dtype=np.complex128, order="F")

    def _r(self, a):  # complex/real Fortran array -> flat float64 view of the same memory
        return a.reshape(-1, order="F").view(np.float64)

    def _fft(self, a, inverse=False):
        # in place on the Fortran-ordered array: transform its C-ordered transpose view
        v = a.T
        f = sfft.ifftn if inverse else sfft.fftn
        out = f(v, workers=self.threads, overwrite_x=True)
        if out is not v:
            v[...] = out

    def deposit(self, x, y, z, q):
        self.rho.fill(0.0)
        self.lib.port_deposit(len(x), x, y, z, q, self._r(self.rho), _I3(*self.n), _D3(*self.lo), _D3(*self.d))

    def solve_freespace(self, offset=(0.0, 0.0, 0.0), rho=None, efield=None):
        rho = self.rho if rho is None else rho
        efield = self.efield if efield is None else efield
        n, s2 = _I3(*self.n), _I3(*self.s2)
        self.lib.port_embed(self._r(self.crho), self._r(rho), n)
        self._fft(self.crho)
        for ic in (1, 2, 3):
            self.lib.port_green_point(self._r(self.cgrn), s2, _D3(*self.d), self.gamma, ic, _D3(*offset))
            self.lib.port_diff8(self._r(self.temp), self._r(self.cgrn), s2)
            self.lib.port_copy_back(self._r(self.cgrn), self._r(self.temp), s2)
            self._fft(self.cgrn)
            self.lib.port_multiply(self._r(self.temp), self._r(self.crho), self._r(self.cgrn), self.crho.size)
            self._fft(self.temp, inverse=True)
            e = np.zeros(self.n, order="F")
            self.lib.port_extract(self._r(e), self._r(self.temp), n, FPEI)
            efield[:, :, :, ic - 1] = e

    def solve(self, at_cathode=False, max_bounds=None):
        self.solve_freespace()
        if at_cathode:
            rho_img = np.asfortranarray(-self.rho[:, :, ::-1])
            e_img = np.zeros_like(self.efield)
            off_z = 2 * self.lo[2] + (float(max_bounds[2]) - self.lo[2])
            self.solve_freespace((0.0, 0.0, off_z), rho_img, e_img)
            self.efield += e_img

    def interpolate(self, x, y, z):
        ex, ey, ez = (np.empty_like(x) for _ in range(3))
        self.lib.port_interpolate(len(x), x, y, z, self._r(self.efield), _I3(*self.n), _D3(*self.lo), _D3(*self.d), ex, ey, ez)
        return ex, ey, ez

    def timed_step(self, x, y, z, q, at_cathode=False, max_bounds=None):
        t0 = time.perf_counter()
        self.deposit(x, y, z, q)
        t1 = time.perf_counter()
        self.solve(at_cathode, max_bounds)
        t2 = time.perf_counter()
        out = self.interpolate(x, y, z)
        t3 = time.perf_counter()
        return out, {"deposit_s": t1 - t0, "solve_s": t2 - t1, "interpolate_s": t3 - t2}
