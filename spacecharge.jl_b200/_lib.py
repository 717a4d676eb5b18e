"""ctypes binding of include/spacecharge_b200.h.

The product path has no CPU fallback: if the CUDA library is missing or no sm_100 device is
present, importing the symbols works (so that symbol/ABI tests can run on a CPU box) but every
compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SCB_LIB selects an experimental build variant (spacecharge.jl_b200/build.py --tag); default: the product library
LIB_PATH = os.environ.get("SCB_LIB") or os.path.join(HERE, "lib", "libspacecharge_b200.so")

SCB_F32, SCB_F64 = 0, 1
SCB_ORDER_RANDOM, SCB_ORDER_CELL, SCB_ORDER_CELL_TILE, SCB_ORDER_AUTO = 0, 1, 2, 3
SCB_OK = 0
STATUS_NAMES = {0: "SCB_OK", -1: "SCB_ERR_INVALID_ARG", -2: "SCB_ERR_UNSUPPORTED", -3: "SCB_ERR_CUDA",
                -4: "SCB_ERR_NO_DEVICE", -5: "SCB_ERR_ALLOC", -6: "SCB_ERR_COMM"}


class LibraryMissing(RuntimeError):
    pass


class ScbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (STATUS_NAMES.get(code, "?"), code, msg))
        self.code = code
        self.message = msg


class scb_options(C.Structure):
    _fields_ = [("green_cache", C.c_int32), ("deposit_mode", C.c_int32), ("particle_order", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class scb_timing(C.Structure):
    _fields_ = [("deposit_ms", C.c_float), ("solve_ms", C.c_float), ("interpolate_ms", C.c_float),
                ("green_ms", C.c_float), ("pass_ms", C.c_float * 8)]


class scb_particle_strides(C.Structure):
    """Element strides of strided / AoS particle arrays (include/spacecharge_b200.h)."""
    _fields_ = [("x", C.c_int64), ("y", C.c_int64), ("z", C.c_int64), ("q", C.c_int64),
                ("ex", C.c_int64), ("ey", C.c_int64), ("ez", C.c_int64), ("reserved", C.c_int64)]


_I64x3 = C.c_int64 * 3
_F64x3 = C.c_double * 3
_vp = C.c_void_p
_PST = C.POINTER(scb_particle_strides)

# name -> (restype, argtypes); must list every symbol include/spacecharge_b200.h declares
SIGNATURES = {
    "scb_version": (C.c_int, []),
    "scb_create": (C.c_int, [C.c_int, _vp, C.POINTER(scb_options), C.POINTER(_vp)]),
    "scb_destroy": (C.c_int, [_vp]),
    "scb_set_stream": (C.c_int, [_vp, _vp]),
    "scb_sync": (C.c_int, [_vp]),
    "scb_last_error": (C.c_char_p, [_vp]),
    "scb_enable_timing": (C.c_int, [_vp, C.c_int]),
    "scb_get_timing": (C.c_int, [_vp, C.POINTER(scb_timing)]),
    "scb_launch_count": (C.c_int64, [_vp]),
    "scb_clear": (C.c_int, [_vp, _vp, _I64x3, C.c_int]),
    "scb_deposit": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _I64x3, _F64x3, _F64x3, C.c_int]),
    "scb_solve": (C.c_int, [_vp, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _F64x3, C.c_double, C.c_int]),
    "scb_solve_freespace": (C.c_int, [_vp, _vp, _vp, C.c_int, _I64x3, _F64x3, C.c_double, _F64x3]),
    "scb_solve_potential": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _F64x3, C.c_double, C.c_int]),
    "scb_bfield": (C.c_int, [_vp, _vp, _vp, C.c_int, _I64x3, C.c_double]),
    "scb_interpolate_kick": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _I64x3, _F64x3, _F64x3,
                                       _vp, _vp, _vp, C.c_double, C.c_double]),
    "scb_interpolate": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _vp, _vp, _vp]),
    "scb_green": (C.c_int, [_vp, _vp, _I64x3, _F64x3, C.c_double, C.c_int, _F64x3, C.c_int]),
    "scb_bounds": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _F64x3, _F64x3]),
    "scb_cell_index": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_int, _F64x3, _F64x3, _vp, _vp, _vp]),
    "scb_step": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _F64x3,
                           C.c_double, C.c_int, _vp, _vp, _vp]),
    "scb_deposit_strided": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _PST, C.c_int, _vp, C.c_int, _I64x3, _F64x3, _F64x3,
                                      C.c_int]),
    "scb_interpolate_strided": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _PST, C.c_int, _vp, C.c_int, _I64x3, _F64x3, _F64x3,
                                          _vp, _vp, _vp]),
    "scb_interpolate_kick_strided": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _PST, C.c_int, _vp, C.c_int, _I64x3, _F64x3,
                                               _F64x3, _vp, _vp, _vp, C.c_double, C.c_double]),
    "scb_bounds_strided": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _PST, C.c_int, _F64x3, _F64x3]),
    "scb_step_strided": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _PST, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3,
                                   _F64x3, _F64x3, C.c_double, C.c_int, _vp, _vp, _vp]),
    "scb_sort_particles": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_int, _I64x3, _F64x3, _F64x3, _vp]),
    "scb_permute": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.c_int]),
    "scb_set_particle_order": (C.c_int, [_vp, C.c_int]),
    "scb_particle_order_fraction": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, C.c_int, _I64x3, _F64x3, _F64x3,
                                              C.POINTER(C.c_double)]),
    "scb_step_host": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3,
                                _F64x3, C.c_double, C.c_int, _vp, _vp, _vp]),
    "scb_step_host_async": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3,
                                      _F64x3, C.c_double, C.c_int, _vp, _vp, _vp]),
    "scb_step_host_sharded_async": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3,
                                              _F64x3, _F64x3, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "scb_step_host_wait": (C.c_int, [_vp]),
    "scb_drop_green_cache": (C.c_int, [_vp]),
    "scb_workspace_bytes": (C.c_int64, [_vp]),
    "scb_comm_unique_id": (C.c_int, [_vp]),
    "scb_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "scb_comm_destroy": (C.c_int, [_vp]),
    "scb_step_sharded": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _F64x3,
                                   C.c_double, C.c_int, _vp, _vp, _vp]),
    "scb_allreduce_rho": (C.c_int, [_vp, _vp, _I64x3, C.c_int]),
    "scb_solve_sharded": (C.c_int, [_vp, _vp, _vp, C.c_int, _I64x3, _F64x3, _F64x3, _F64x3, C.c_double, C.c_int]),
    # include/spacecharge_b200_debug.h
    "scb_debug_fft_lines": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double]),
    "scb_debug_fft_x_r2c": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "scb_debug_l2_probe": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_double)]),
    "scb_debug_fft_x_c2r": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double]),
}

_lib = None


def load():
    """dlopen the in-tree shared object and attach the prototypes.  Raises LibraryMissing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "CUDA extension %s not built (run `python __graft_entry__.py build`); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def i64x3(v):
    return _I64x3(int(v[0]), int(v[1]), int(v[2]))


def f64x3(v):
    return _F64x3(float(v[0]), float(v[1]), float(v[2]))
