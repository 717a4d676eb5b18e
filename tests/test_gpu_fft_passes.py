"""Every supported length of the hand-written FFT engine, pass by pass, against numpy.fft
(through the debug hooks of include/spacecharge_b200_debug.h)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LENGTHS = [8, 16, 32, 64, 128, 256, 512, 1024, 2048]
DT = {"f32": (0, np.float32, np.complex64, 2e-6), "f64": (1, np.float64, np.complex128, 5e-15)}


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("N", LENGTHS)
def test_strided_pass_forward_padded_and_inverse_pruned(scb, prec, N):
    import torch
    tag, rt, ct, tol = DT[prec]
    hd = scb.default_handle()
    rng = np.random.default_rng(N)
    ninner, nouter = 11, 3
    n_in = N // 2 if N > 8 else 3
    # forward: [o][pos][i], pos < n_in, zero padded to N
    a = (rng.standard_normal((nouter, n_in, ninner)) + 1j * rng.standard_normal((nouter, n_in, ninner))).astype(ct)
    d_in = _dev(torch, a.view(rt))
    d_out = torch.zeros((nouter, N, ninner, 2), dtype=d_in.dtype, device="cuda")
    hd.check(hd.lib.scb_debug_fft_lines(hd.h, tag, N, -1, d_in.data_ptr(), d_out.data_ptr(), n_in, N, ninner,
                                        ninner, n_in * ninner, ninner, N * ninner, nouter, 1.0))
    got = d_out.cpu().numpy().view(ct)[..., 0]
    want = np.fft.fft(a.astype(np.complex128), n=N, axis=1)
    assert _rel(got, want) < tol * np.log2(N)
    # inverse, keep the first n_out bins, scaled
    n_out = n_in
    b = (rng.standard_normal((nouter, N, ninner)) + 1j * rng.standard_normal((nouter, N, ninner))).astype(ct)
    d_in = _dev(torch, b.view(rt))
    d_out = torch.zeros((nouter, n_out, ninner, 2), dtype=d_in.dtype, device="cuda")
    hd.check(hd.lib.scb_debug_fft_lines(hd.h, tag, N, +1, d_in.data_ptr(), d_out.data_ptr(), N, n_out, ninner,
                                        ninner, N * ninner, ninner, n_out * ninner, nouter, 0.5))
    got = d_out.cpu().numpy().view(ct)[..., 0]
    want = 0.5 * N * np.fft.ifft(b.astype(np.complex128), axis=1)[:, :n_out, :]
    assert _rel(got, want) < tol * np.log2(N)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("N", LENGTHS)
def test_x_passes_real_to_half_complex_and_back(scb, prec, N):
    import torch
    tag, rt, ct, tol = DT[prec]
    hd = scb.default_handle()
    rng = np.random.default_rng(1000 + N)
    nlines = 7  # odd: the last line is paired with zeros
    n_real = N // 2 if N > 8 else 3
    PX = (N // 2 + 1 + 7) // 8 * 8
    a = rng.standard_normal((nlines, n_real)).astype(rt)
    d_in = _dev(torch, a)
    d_out = torch.zeros((nlines, PX, 2), dtype=d_in.dtype, device="cuda")
    hd.check(hd.lib.scb_debug_fft_x_r2c(hd.h, tag, N, d_in.data_ptr(), d_out.data_ptr(), nlines, n_real, n_real, PX))
    got = d_out.cpu().numpy().view(ct)[..., 0][:, :N // 2 + 1]
    want = np.fft.rfft(a.astype(np.float64), n=N, axis=1)
    assert _rel(got, want) < tol * np.log2(N)
    # back: Hermitian half spectra of real lines -> first n_real samples
    full = rng.standard_normal((nlines, N))
    spec = np.zeros((nlines, PX), dtype=ct)
    spec[:, :N // 2 + 1] = np.fft.rfft(full, axis=1)
    d_in = _dev(torch, spec.view(rt))
    d_out = torch.zeros((nlines, n_real), dtype=d_in.dtype, device="cuda")
    hd.check(hd.lib.scb_debug_fft_x_c2r(hd.h, tag, N, d_in.data_ptr(), d_out.data_ptr(), nlines, n_real, n_real, PX, 1.0 / N))
    got = d_out.cpu().numpy()
    assert _rel(got, full[:, :n_real]) < tol * np.log2(N) * 4
