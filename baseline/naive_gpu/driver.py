"""ctypes driver of the secondary GPU baseline (reference structure in plain CUDA + cuFFT).
Bench/test infrastructure only; Float64, default CUDA stream."""
import ctypes as C
import os
import importlib.util

HERE = os.path.dirname(os.path.abspath(__file__))


def _build():
    spec = importlib.util.spec_from_file_location("naive_gpu_build", os.path.join(HERE, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


class RefGpu:
    def __init__(self):
        self.lib = C.CDLL(_build())
        I3, D3, vp = C.c_int * 3, C.c_double * 3, C.c_void_p
        self.I3, self.D3 = I3, D3
        self.lib.refgpu_deposit.argtypes = [C.c_longlong, vp, vp, vp, vp, vp, I3, D3, D3]
        self.lib.refgpu_solve.argtypes = [vp, vp, I3, D3, D3, D3, C.c_double, C.c_int]
        self.lib.refgpu_interpolate.argtypes = [C.c_longlong, vp, vp, vp, vp, I3, D3, D3, vp, vp, vp]

    def step(self, grid, lo, hi, delta, gamma, at_cathode, x, y, z, q, rho, efield, ex, ey, ez, events=None):
        """torch CUDA tensors (float64, contiguous).  rho/efield in the reference layout (x fastest)."""
        n, lo3, hi3, d3 = self.I3(*grid), self.D3(*lo), self.D3(*hi), self.D3(*delta)
        rec = (lambda i: events[i].record()) if events else (lambda i: None)
        rec(0)
        rc = self.lib.refgpu_deposit(x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), q.data_ptr(), rho.data_ptr(), n, lo3, d3)
        rec(1)
        rc |= self.lib.refgpu_solve(rho.data_ptr(), efield.data_ptr(), n, lo3, hi3, d3, float(gamma), 1 if at_cathode else 0)
        rec(2)
        rc |= self.lib.refgpu_interpolate(x.numel(), x.data_ptr(), y.data_ptr(), z.data_ptr(), efield.data_ptr(), n, lo3, d3,
                                          ex.data_ptr(), ey.data_ptr(), ez.data_ptr())
        rec(3)
        if rc != 0:
            raise RuntimeError("naive GPU baseline failed with CUDA error %d" % rc)
