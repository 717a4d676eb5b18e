"""NumPy model of the *GPU* algorithm (not of the reference): wrap-around IGF placement on
power-of-two padded lengths, real-to-complex transforms, one inverse per component and the
cathode image folded into the spectral multiply.  Used by the CPU tests to show that the
restructured algorithm equals the oracle's reference-structured one to round-off."""
import numpy as np
from oracle import spacecharge_oracle as so


def padded_len(n):
    L = 8
    while L < 2 * n:
        L *= 2
    return L


def igf_block(n, delta, gamma, icomp, offset=(0.0, 0.0, 0.0)):
    """IGF(d) for d in [-(n-1), n-1]^3, float64, via point-wise corners + differencing."""
    nx, ny, nz = n
    g = so.get_green_function((2 * nx, 2 * ny, 2 * nz), delta, gamma, icomp, offset, np.float64)
    return g[:-1, :-1, :-1]  # index i <-> d = i - (n-1)


def green_spectra(n, delta, gamma, offset_z=None):
    nx, ny, nz = n
    L = tuple(padded_len(v) for v in n)
    out = []
    for ic in (1, 2, 3):
        if offset_z is None:
            blk = igf_block(n, delta, gamma, ic)
            g = np.zeros(L)
            ix = (np.arange(2 * nx - 1) - (nx - 1)) % L[0]
            iy = (np.arange(2 * ny - 1) - (ny - 1)) % L[1]
            iz = (np.arange(2 * nz - 1) - (nz - 1)) % L[2]
            g[np.ix_(ix, iy, iz)] = blk
        else:
            blk = igf_block(n, delta, gamma, ic, (0.0, 0.0, offset_z))
            g = np.zeros(L)
            ix = (np.arange(2 * nx - 1) - (nx - 1)) % L[0]
            iy = (np.arange(2 * ny - 1) - (ny - 1)) % L[1]
            iz = np.arange(2 * nz - 1)  # s = dz + (nz-1): correlation placement
            g[np.ix_(ix, iy, iz)] = -blk
        out.append(np.fft.rfftn(g, axes=(2, 1, 0)))  # half spectrum along x
    return out


def solve_fused(rho, delta, gamma, min_z=None, max_z=None, at_cathode=False):
    n = rho.shape
    L = tuple(padded_len(v) for v in n)
    pad = np.zeros(L)
    pad[:n[0], :n[1], :n[2]] = rho
    R = np.fft.rfftn(pad, axes=(2, 1, 0))
    G = green_spectra(n, delta, gamma)
    if at_cathode:
        H = green_spectra(n, delta, gamma, offset_z=min_z + max_z)
        Rm = R[:, :, (-np.arange(L[2])) % L[2]]
    E = np.zeros(n + (3,), order="F")
    for c in range(3):
        S = R * G[c]
        if at_cathode:
            S = S + Rm * H[c]
        e = np.fft.irfftn(S, s=(L[2], L[1], L[0]), axes=(2, 1, 0))
        E[..., c] = so.FPEI * e[:n[0], :n[1], :n[2]]
    return E
