"""Thread-level NumPy model of the even/odd-bin z pass (k_z_eo in csrc/fft_passes.cuh), N = 512.

A zero-padded line x[0..255] -> X = FFT512(pad(x)) is never formed as one transform: the even bins are
FFT256(x), the odd bins FFT256(x * w512^n); each 256-point transform runs on 16 threads x 16 values as
radix-16 x radix-16 with ONE exchange.  The inverse keeps only y[0..255] = IFFT256(W_even) + w512^-n * IFFT256(W_odd).
Half-warp h = 0 carries the even bins, h = 1 the odd ones.  This file checks the index algebra and the codelets
against numpy.fft; run it directly."""
import numpy as np


def dft4(x0, x1, x2, x3, d):          # d = -1 forward, +1 inverse; natural order in and out
    q = (lambda a: a * (-1j)) if d < 0 else (lambda a: a * 1j)
    a0, a1, a2, a3 = x0 + x2, x0 - x2, x1 + x3, q(x1 - x3)
    return a0 + a2, a1 + a3, a0 - a2, a1 - a3


def dft16(v, d):
    """v: list of 16 arrays, natural order in, natural order out (4 x 4, twiddles w16^(q0*k0))."""
    v = list(v)
    w16 = np.exp(d * 2j * np.pi / 16)
    for q0 in range(4):
        v[q0], v[q0 + 4], v[q0 + 8], v[q0 + 12] = dft4(v[q0], v[q0 + 4], v[q0 + 8], v[q0 + 12], d)
    for q0 in range(1, 4):
        for k0 in range(1, 4):
            v[q0 + 4 * k0] = v[q0 + 4 * k0] * w16 ** (q0 * k0)
    for k0 in range(4):
        v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3] = dft4(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3], d)
    out = [None] * 16
    for k0 in range(4):
        for k1 in range(4):
            out[k0 + 4 * k1] = v[4 * k0 + k1]
    return out


def forward(x):
    """x: (256,) -> spec[h][t][k2] = bin 2*(t + 16*k2) + h of the 512-point transform of the padded line."""
    tw = np.exp(-2j * np.pi * np.arange(512) / 512)
    t = np.arange(16)
    spec = np.zeros((2, 16, 16), complex)
    for h in range(2):
        v = [x[t + 16 * q] * (np.exp(-2j * np.pi * q / 32) if h else 1.0) for q in range(16)]
        a = dft16(v, -1)
        p0 = tw[t] if h else np.ones(16)
        b = tw[2 * t]
        a = [a[k1] * p0 * b ** k1 for k1 in range(16)]
        ex = np.zeros((16, 17), complex)            # exchange buffer [k1][t], pitch 17
        for k1 in range(16):
            ex[k1, t] = a[k1]
        a2 = [ex[t, t2] for t2 in range(16)]        # thread t now plays k1 = t
        s = dft16(a2, -1)
        for k2 in range(16):
            spec[h, :, k2] = s[k2]
    return spec


def inverse(wspec):
    """wspec[h][t][k2] -> y[0..255] (unnormalised inverse 512-point transform, first half)."""
    tw = np.exp(-2j * np.pi * np.arange(512) / 512)
    t = np.arange(16)
    parts = []
    for h in range(2):
        a = dft16([wspec[h, :, k2] for k2 in range(16)], +1)
        b = np.conj(tw[2 * t + h])
        a = [a[n1] * b ** n1 for n1 in range(16)]
        ex = np.zeros((16, 17), complex)
        for n1 in range(16):
            ex[n1, t] = a[n1]
        a2 = [ex[t, t2] for t2 in range(16)]
        y = dft16(a2, +1)                           # y[n2] at thread n1 = t: n = t + 16*n2
        if h:
            y = [y[n2] * np.exp(2j * np.pi * n2 / 32) for n2 in range(16)]
        parts.append(y)
    out = np.zeros(256, complex)
    for n2 in range(16):
        out[t + 16 * n2] = parts[0][n2] + parts[1][n2]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    X = np.fft.fft(np.concatenate([x, np.zeros(256)]))
    spec = forward(x)
    t = np.arange(16)
    err = 0.0
    for h in range(2):
        for k2 in range(16):
            err = max(err, np.abs(spec[h, :, k2] - X[2 * (t + 16 * k2) + h]).max())
    print("forward max err", err)
    g = rng.standard_normal(512) + 1j * rng.standard_normal(512)
    W = X * g
    yref = np.fft.ifft(W)[:256] * 512
    wspec = np.zeros_like(spec)
    for h in range(2):
        for k2 in range(16):
            wspec[h, :, k2] = spec[h, :, k2] * g[2 * (t + 16 * k2) + h]
    y = inverse(wspec)
    print("inverse max err", np.abs(y - yref).max() / np.abs(yref).max())
    assert err < 1e-10 and np.abs(y - yref).max() / np.abs(yref).max() < 1e-12
