"""NumPy model of the kx-slab decomposition of the solve (run_solve_sharded_kx in csrc/api.cu): G "ranks" hold flat
buffers with exactly the layouts and index arithmetic of the CUDA path, so that the host-side contract of the two fused
exchanges can be checked on a CPU box:

  F1  (k_x_r2c, XParams::split / line0 / out_peer): bin k of line l of rank `me` goes to rank k // PXl at
      A1[(l + me*ny*nzl) * PXl + k % PXl]                                      -> A1 = [z_global][y][kx_l]
  F2, z pass, B2 local on [z][ky][kx_l] (pitch PXl); the Green spectrum is indexed at kx0 + kx_l
  B2  (k_lines, use_peers == 2 / out_osplit): plane z of rank `me` goes to rank z // nzl at
      DR[c*G*blk + me*blk + ((z % nzl)*ny + y) * PXl + kx_l]                    -> DR = [c][src][z_l][y][kx_l]
  B3  (k_x_c2r, XParams::split / sblock): bin k of line l is read from DR[c*G*blk + (k // PXl)*blk + l*PXl + k % PXl]

The arithmetic itself (FFTs, Green multiply) is numpy's; what is modelled is WHERE every number lives."""
import numpy as np

from fused_model import green_spectra, padded_len
from oracle import spacecharge_oracle as so


def solve_kx_slabs(rho, delta, gamma, G):
    """free-space E (nx, ny, nz, 3) computed by G model ranks; rho: (nx, ny, nz) full grid (already summed)"""
    nx, ny, nz = rho.shape
    L = tuple(padded_len(v) for v in rho.shape)
    ninner = L[0] // 2 + 1
    PX = (ninner + 7) // 8 * 8
    assert nz % G == 0 and PX % G == 0
    nzl, PXl = nz // G, PX // G
    blk = PXl * ny * nzl
    a1 = PXl * ny * nz
    spectra = green_spectra(rho.shape, delta, gamma)            # [c][kx][ky][kz], half spectrum along x
    A1 = [np.zeros(a1, dtype=complex) for _ in range(G)]
    DR = [np.zeros(3 * a1, dtype=complex) for _ in range(G)]
    # ---- F1 on every rank's z slab, bins stored to the owner of their kx block
    for me in range(G):
        line0 = me * ny * nzl
        for zl in range(nzl):
            for y in range(ny):
                line = y + ny * zl
                real = np.zeros(L[0])
                real[:nx] = rho[:, y, me * nzl + zl]
                spec = np.fft.rfft(real)                         # bins 0 .. L/2
                for k in range(ninner):
                    A1[k // PXl][(line + line0) * PXl + k % PXl] = spec[k]
    # ---- F2, z pass, B2 on every rank's kx slab; B2's planes stored to the owners of their z slabs
    for me in range(G):
        kx0 = me * PXl
        ninl = max(0, min(PXl, ninner - kx0))
        if ninl == 0:
            continue
        a = A1[me].reshape(nz, ny, PXl)[:, :, :ninl]             # [z][y][kx_l]
        pad_y = np.zeros((nz, L[1], ninl), dtype=complex)
        pad_y[:, :ny, :] = a
        b = np.fft.fft(pad_y, axis=1)                            # F2: [z][ky][kx_l]
        pad_z = np.zeros((L[2], L[1], ninl), dtype=complex)
        pad_z[:nz] = b
        bz = np.fft.fft(pad_z, axis=0)                           # forward along z
        for c in range(3):
            g = spectra[c][kx0:kx0 + ninl].transpose(2, 1, 0)    # Green spectrum at the GLOBAL kx, as [kz][ky][kx_l]
            cz = np.fft.ifft(bz * g, axis=0)[:nz]                # inverse along z, first nz kept
            d = np.fft.ifft(cz, axis=1)[:, :ny, :]               # B2: inverse along y, first ny kept -> [z][y][kx_l]
            for z in range(nz):
                dst = DR[z // nzl]
                base = c * G * blk + me * blk
                for y in range(ny):
                    off = base + ((z % nzl) * ny + y) * PXl
                    dst[off:off + ninl] = d[z, y, :]
    # ---- B3 on every rank's z slab, kx gathered from the ranks' blocks
    E = np.zeros((nx, ny, nz, 3))
    for me in range(G):
        for c in range(3):
            for zl in range(nzl):
                for y in range(ny):
                    line = y + ny * zl
                    spec = np.array([DR[me][c * G * blk + (k // PXl) * blk + line * PXl + k % PXl] for k in range(ninner)])
                    E[:, y, me * nzl + zl, c] = so.FPEI * np.fft.irfft(spec, n=L[0])[:nx]
    return E
