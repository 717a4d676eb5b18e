// fft_engine.cuh -- in-register / shared-memory Stockham FFT engine (power-of-two lengths 8..2048).
//
// One line of length N is transformed by TPL = N/8 threads.  Thread j owns 8 complex values:
// on entry v[q] is the element at position j + q*TPL, on exit v[q] is the bin j + q*TPL (both
// natural order).  Stages are radix-8 while the remaining factor allows, then one radix-4 or
// radix-2 stage; between stages the values cross threads through a shared-memory exchange
// buffer described by a layout functor.  tools/fft_engine_model.py is the NumPy model of exactly
// this schedule (checked against numpy.fft for every supported N).
//
// No cuFFTDx in this image: this replaces it.
#pragma once
#include "common.cuh"

namespace scb {

// ---- small DFTs, natural order in / natural order out -----------------------------------
template <int DIR, typename C> __device__ __forceinline__ void dft2(C& a, C& b) {
    C t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <int DIR, typename C> __device__ __forceinline__ void dft4(C& x0, C& x1, C& x2, C& x3) {
    C a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3);
    C a3 = cmul_qturn<DIR>(csub(x1, x3));
    x0 = cadd(a0, a2);
    x2 = csub(a0, a2);
    x1 = cadd(a1, a3);
    x3 = csub(a1, a3);
}

template <int DIR, typename C>
__device__ __forceinline__ void dft8(C& x0, C& x1, C& x2, C& x3, C& x4, C& x5, C& x6, C& x7) {
    using R = typename real_of<C>::type;
    const R h = (R)0.70710678118654752440;
    // even / odd halves
    dft4<DIR>(x0, x2, x4, x6);  // E0..E3 in x0,x2,x4,x6
    dft4<DIR>(x1, x3, x5, x7);  // O0..O3 in x1,x3,x5,x7
    // O1 *= W8^1, O2 *= W8^2, O3 *= W8^3  (forward W8 = exp(-i pi/4); inverse: conjugate)
    C o1, o2, o3;
    if (DIR < 0) {
        o1 = cmake<C>((x3.x + x3.y) * h, (x3.y - x3.x) * h);
        o3 = cmake<C>((x7.y - x7.x) * h, -(x7.x + x7.y) * h);
    } else {
        o1 = cmake<C>((x3.x - x3.y) * h, (x3.x + x3.y) * h);
        o3 = cmake<C>(-(x7.x + x7.y) * h, (x7.x - x7.y) * h);
    }
    o2 = cmul_qturn<DIR>(x5);
    C e0 = x0, e1 = x2, e2 = x4, e3 = x6, o0 = x1;
    x0 = cadd(e0, o0);
    x4 = csub(e0, o0);
    x1 = cadd(e1, o1);
    x5 = csub(e1, o1);
    x2 = cadd(e2, o2);
    x6 = csub(e2, o2);
    x3 = cadd(e3, o3);
    x7 = csub(e3, o3);
}

// ---- shared-memory exchange layouts ------------------------------------------------------
// "line-in-lockstep" layout for the strided (y / z) passes: TX lines advance together, the
// TX values of one position form one row.  Rows of >= 128 bytes are bank-conflict free for any
// position pattern; narrower rows get one padding row per 8 rows.
template <typename C, int TX>
struct LayoutRows {
    static constexpr bool PAD = (sizeof(C) * TX < 128);
    C* buf;
    int tx;
    __device__ __forceinline__ LayoutRows(C* b, int t) : buf(b), tx(t) {}
    __device__ __forceinline__ int idx(int pos) const { return (PAD ? pos + (pos >> 3) : pos) * TX + tx; }
    __device__ __forceinline__ void st(int pos, C v) const { buf[idx(pos)] = v; }
    __device__ __forceinline__ C ld(int pos) const { return buf[idx(pos)]; }
    __host__ __device__ static constexpr int rows(int N) { return PAD ? N + N / 8 : N; }
    __host__ __device__ static constexpr size_t bytes(int N) { return (size_t)rows(N) * TX * sizeof(C); }
};

// contiguous-line layout for the x passes: each line is a padded row of its own.
template <typename C>
struct LayoutLine {
    C* buf;  // already offset to this thread's line
    __device__ __forceinline__ explicit LayoutLine(C* b) : buf(b) {}
    __device__ __forceinline__ int idx(int pos) const { return pos + (pos >> 3); }
    __device__ __forceinline__ void st(int pos, C v) const { buf[idx(pos)] = v; }
    __device__ __forceinline__ C ld(int pos) const { return buf[idx(pos)]; }
    __host__ __device__ static constexpr int row(int N) { return N + N / 8 + 1; }
};

// ---- per-thread twiddle registers --------------------------------------------------------------
// The twiddles a thread needs depend only on its position j in the line, not on the data: stage Ns uses
// w1 = root^t, w2 = root^2t, w4 = root^4t (radix 8) with t = (jp mod Ns) * N/(Ns*R).  ncu on the fused z pass
// showed these table loads as the largest stall reason (long scoreboard) and ~18 % of the LSU wavefronts, so only
// w1 is fetched -- once per thread, before the first transform, and kept in registers across all transforms of
// the kernel -- and w2 = w1^2, w4 = w2^2 are formed in registers (a few ulp, far below the parity bar).
template <int N, int Ns = 1> __host__ __device__ constexpr int tw_count() {
    constexpr int REM = N / Ns;
    constexpr int R = REM >= 8 ? 8 : REM;
    constexpr int G = 8 / R;
    if constexpr (Ns * R < N) return (Ns > 1 ? G : 0) + tw_count<N, Ns * R>();
    else return Ns > 1 ? G : 0;
}

template <typename T, int N, int Ns = 1, int OFF = 0>
__device__ __forceinline__ void fft_twiddles(cx_t<T> (&w)[tw_count<N>() > 0 ? tw_count<N>() : 1], const int j,
                                             const cx_t<T>* __restrict__ tw) {
    constexpr int TPL = N / 8;
    constexpr int REM = N / Ns;
    constexpr int R = REM >= 8 ? 8 : REM;
    constexpr int G = 8 / R;
    if constexpr (Ns > 1) {
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int jp = j + m * TPL;
            w[OFF + m] = __ldg(tw + (jp & (Ns - 1)) * (N / (Ns * R)));
        }
    }
    if constexpr (Ns * R < N) fft_twiddles<T, N, Ns * R, OFF + (Ns > 1 ? G : 0)>(w, j, tw);
}

// ---- one Stockham stage (recursive over Ns) ----------------------------------------------
template <typename T, int N, int DIR, int Ns, int OFF, typename Lay>
__device__ __forceinline__ void fft_stage(cx_t<T> (&v)[8], const Lay& lay, const int j,
                                          const cx_t<T> (&wr)[tw_count<N>() > 0 ? tw_count<N>() : 1]) {
    using C = cx_t<T>;
    constexpr int TPL = N / 8;
    constexpr int REM = N / Ns;
    constexpr int R = REM >= 8 ? 8 : REM;  // 8, 4 or 2
    constexpr int G = 8 / R;               // butterflies per thread in this stage
    static_assert(R == 8 || R == 4 || R == 2, "bad radix");

#pragma unroll
    for (int m = 0; m < G; ++m) {
        if constexpr (Ns > 1) {
            const C w1 = tw_dir<DIR>(wr[OFF + m]);
            if constexpr (R == 8) {
                const C w2 = cmul(w1, w1);
                const C w4 = cmul(w2, w2);
                const C w3 = cmul(w1, w2);
                v[m + 1 * G] = cmul(v[m + 1 * G], w1);
                v[m + 2 * G] = cmul(v[m + 2 * G], w2);
                v[m + 3 * G] = cmul(v[m + 3 * G], w3);
                v[m + 4 * G] = cmul(v[m + 4 * G], w4);
                v[m + 5 * G] = cmul(v[m + 5 * G], cmul(w4, w1));
                v[m + 6 * G] = cmul(v[m + 6 * G], cmul(w4, w2));
                v[m + 7 * G] = cmul(v[m + 7 * G], cmul(w4, w3));
            } else if constexpr (R == 4) {
                const C w2 = cmul(w1, w1);
                v[m + 1 * G] = cmul(v[m + 1 * G], w1);
                v[m + 2 * G] = cmul(v[m + 2 * G], w2);
                v[m + 3 * G] = cmul(v[m + 3 * G], cmul(w1, w2));
            } else {
                v[m + 1 * G] = cmul(v[m + 1 * G], w1);
            }
        }
        if constexpr (R == 8) {
            dft8<DIR>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        } else if constexpr (R == 4) {
            dft4<DIR>(v[m], v[m + G], v[m + 2 * G], v[m + 3 * G]);
        } else {
            dft2<DIR>(v[m], v[m + G]);
        }
    }

    if constexpr (Ns * R < N) {
        __syncthreads();  // everyone is done reading the previous contents of the buffer
#pragma unroll
        for (int m = 0; m < G; ++m) {
            const int jp = j + m * TPL;
            const int base = (jp / Ns) * (Ns * R) + (jp & (Ns - 1));
#pragma unroll
            for (int r = 0; r < R; ++r) lay.st(base + r * Ns, v[m + r * G]);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = lay.ld(j + q * TPL);
        fft_stage<T, N, DIR, Ns * R, OFF + (Ns > 1 ? G : 0), Lay>(v, lay, j, wr);
    }
}

// number of twiddle registers (complex) a thread keeps for length N
template <int N> struct TwN { static constexpr int value = tw_count<N>() > 0 ? tw_count<N>() : 1; };

// Full transform of the line owned by this thread group, twiddles already in registers.  Unnormalised.
template <typename T, int N, int DIR, typename Lay>
__device__ __forceinline__ void fft_line(cx_t<T> (&v)[8], const Lay& lay, const int j,
                                         const cx_t<T> (&wr)[TwN<N>::value]) {
    fft_stage<T, N, DIR, 1, 0, Lay>(v, lay, j, wr);
}

// Convenience for kernels that run a single transform: fetch the twiddles, then transform.
template <typename T, int N, int DIR, typename Lay>
__device__ __forceinline__ void fft_line(cx_t<T> (&v)[8], const Lay& lay, const int j,
                                         const cx_t<T>* __restrict__ tw) {
    cx_t<T> wr[TwN<N>::value];
    fft_twiddles<T, N>(wr, j, tw);
    fft_stage<T, N, DIR, 1, 0, Lay>(v, lay, j, wr);
}

}  // namespace scb
