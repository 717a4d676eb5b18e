/* spacecharge_b200_debug.h -- parity hooks for the individual FFT passes.
 *
 * NOT part of the drop-in surface (the reference has no such entry points: its FFTs are
 * AbstractFFTs plans, src/mesh.jl:60-61).  The GPU tests use these to check every supported
 * transform length of the hand-written engine (csrc/fft_engine.cuh) against numpy.fft, pass by
 * pass, before the passes are composed into scb_solve.  Same conventions as
 * spacecharge_b200.h: device pointers, async on the handle's stream, int status codes.
 * Complex arrays are interleaved (re, im) of the given dtype.
 */
#ifndef SPACECHARGE_B200_DEBUG_H
#define SPACECHARGE_B200_DEBUG_H

#include "spacecharge_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Strided complex pass: for o in [0,nouter), i in [0,ninner): line elements
 * in[i + pos*in_sline + o*in_souter], pos < n_in (zero beyond), unnormalised DFT of length N in
 * direction dir (-1 forward, +1 inverse), bins < n_out stored to
 * out[i + pos*out_sline + o*out_souter] times scale. */
SCB_API int scb_debug_fft_lines(scb_handle* h, int dt, int N, int dir, const void* in, void* out,
                                int n_in, int n_out, int ninner, int64_t in_sline, int64_t in_souter,
                                int64_t out_sline, int64_t out_souter, int nouter, double scale);

/* Real lines (n_real valid reals each, stride real_sline, zero-padded to N) -> half spectra
 * out[k + PX*line], k in [0, N/2]. */
SCB_API int scb_debug_fft_x_r2c(scb_handle* h, int dt, int N, const void* in_real, void* out_cplx,
                                int64_t nlines, int64_t real_sline, int n_real, int PX);

/* Half spectra in[k + PX*line] -> first n_real samples of the length-N inverse real transform
 * (unnormalised) times scale. */
SCB_API int scb_debug_fft_x_c2r(scb_handle* h, int dt, int N, const void* in_cplx, void* out_real,
                                int64_t nlines, int64_t real_sline, int n_real, int PX, double scale);

/* Measurement only (bench.py): ceiling of the scattered 32-byte sector traffic that bounds the particle passes.
 * mode 0: lane pairs issue the gather's 256-bit no-allocate loads at two adjacent pseudo-random records of a zeroed buffer of
 * buffer_bytes; mode 1: every four-lane group issues the tile deposit's fp64 reduction into a random 32-byte tile.
 * Nothing else runs in the kernel.  Synchronous; returns 32-byte sector operations per second (best of three warm
 * repetitions, CUDA events).  A buffer smaller than the L2 measures the L2 itself, a larger one adds DRAM misses. */
SCB_API int scb_debug_l2_probe(scb_handle* h, int mode, int64_t buffer_bytes, int iters, double* sector_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
