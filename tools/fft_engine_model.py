"""Thread-level NumPy model of the device FFT engine in csrc/fft_engine.cuh.

Each "thread" j of a line owns 8 elements at positions j + q*TPL (q=0..7) on entry and the bins
j + q*TPL on exit; stages are Stockham radix-8/4/2 with a shared-memory exchange in between.
Run as a script to check every supported length and both directions against numpy.fft."""
import numpy as np


def dft_small(x, DIR):
    R = len(x)
    k = np.arange(R)
    W = np.exp(DIR * 2j * np.pi * np.outer(k, k) / R)
    return W @ x


def engine(line, DIR):
    N = len(line)
    TPL = N // 8
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    if DIR > 0:
        tw = tw.conj()
    v = np.zeros((TPL, 8), dtype=complex)
    for j in range(TPL):
        for q in range(8):
            v[j, q] = line[j + q * TPL]
    Ns = 1
    while Ns < N:
        rem = N // Ns
        R = 8 if rem >= 8 else rem
        G = 8 // R
        smem = np.zeros(N, dtype=complex)
        for j in range(TPL):
            for m in range(G):
                jp = j + m * TPL
                idx = [m + r * G for r in range(R)]
                x = v[j, idx].copy()
                if Ns > 1:
                    k = jp % Ns
                    t = k * (N // (Ns * R))
                    for r in range(1, R):
                        x[r] *= tw[t * r]
                x = dft_small(x, DIR)
                v[j, idx] = x
                if Ns * R < N:
                    base = (jp // Ns) * Ns * R + (jp % Ns)
                    for r in range(R):
                        smem[base + r * Ns] = x[r]
        if Ns * R < N:
            for j in range(TPL):
                for q in range(8):
                    v[j, q] = smem[j + q * TPL]
        Ns *= R
    out = np.zeros(N, dtype=complex)
    for j in range(TPL):
        for q in range(8):
            out[j + q * TPL] = v[j, q]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for N in (8, 16, 32, 64, 128, 256, 512, 1024, 2048):
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        f = engine(x, -1)
        b = engine(x, +1)
        print(N, np.abs(f - np.fft.fft(x)).max(), np.abs(b - np.fft.ifft(x) * N).max())
