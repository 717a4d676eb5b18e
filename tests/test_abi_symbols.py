"""The C-ABI library loads on a CPU box and exports every symbol include/*.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("spacecharge_b200.h", "spacecharge_b200_debug.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        names |= set(re.findall(r"SCB_API[^;]*?\b(scb_\w+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_reference_surface():
    names = declared_symbols()
    # one entry point per reference function on the hot path
    for required in ("scb_clear", "scb_deposit", "scb_solve", "scb_solve_freespace", "scb_interpolate", "scb_green",
                     "scb_bounds", "scb_create", "scb_destroy", "scb_last_error", "scb_step", "scb_step_host", "scb_step_host_async", "scb_step_host_wait"):
        assert required in names
    assert len(names) >= 24


def test_library_exports_every_declared_symbol(scb):
    lib_path = scb._lib.LIB_PATH
    assert os.path.exists(lib_path), "build the extension first: python __graft_entry__.py build"
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    # and the Python binding knows a prototype for each of them
    assert set(declared_symbols()) == set(scb._lib.SIGNATURES)
    assert scb._lib.load().scb_version() >= 100


def test_no_cpu_fallback(scb):
    """Without a GPU the product must fail loudly, never fall back to the oracle or any CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises((scb.LibraryMissing, scb.ScbError)):
        scb.Handle(0)
    src = open(os.path.join(ROOT, "spacecharge.jl_b200", "__init__.py")).read() + \
        open(os.path.join(ROOT, "spacecharge.jl_b200", "_lib.py")).read()
    assert "oracle" not in src.replace("the oracle", "")


def test_ctypes_prototypes_match_the_header(scb):
    """Argument count and kind (pointer / int64 / int / double) of every ctypes prototype against the C declaration."""
    import ctypes as C
    from test_julia_shim_signatures import header_prototypes
    _lib = scb._lib

    def kind(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C.Array):
            return "ptr"
        return {C.c_int64: "i64", C.c_int: "i32", C.c_double: "f64"}[t]

    protos = header_prototypes()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert [kind(t) for t in argtypes] == protos[name], name
