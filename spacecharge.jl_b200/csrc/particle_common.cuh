// particle_common.cuh -- device helpers shared by the particle kernels (particles.cu, sorted.cu): promotion rule,
// cell location with the reference's arithmetic, L2 residency hints.
//
// Arithmetic follows the reference exactly: normalised coordinate with a true IEEE division
// in promote(P, T), floor, fraction.  Translation units that include this header are compiled with -fmad=false.
#pragma once
#include "kernels.h"

namespace scb {

template <typename P, typename T> struct promote { using type = double; };
template <> struct promote<float, float> { using type = float; };

// L2 residency policy for the grid-side accesses.  The particle streams (3.2 GB in, 2.4 GB out per pass) are
// read/written once with evict-first loads/stores, but they still wash the scattered grid records out of the
// 126 MB L2: ncu showed only 70 % of the gather's sector reads and 42 % of the deposit's L2 look-ups hitting although
// the +-3 sigma core that 99 % of the particles touch is < 100 MB.  Grid records / accumulator tiles are therefore
// accessed with an evict_last policy so that the streams are evicted in preference to them.
__device__ __forceinline__ unsigned long long l2_policy(int keep) {
    unsigned long long p;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void red_add_hint(double* a, double v, unsigned long long pol) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_hint(float* a, float v, unsigned long long pol) {
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}

// index of particle i in an array with element stride s; ST = false (contiguous callers) compiles to i itself, so the
// tuned kernels are unchanged for the reference's SoA layout
template <bool ST> __device__ __forceinline__ long long pidx(long long i, long long s) { return ST ? i * s : i; }

template <typename W> struct CellW {
    int i[3];
    W f[3];
};

// t = a / d, correctly rounded, without the division sequence.  rinv = RN(1/d) comes from the host (a true IEEE
// division, make_geom).  Markstein's correction steps: with y = RN(1/d) and q a faithful approximation of a/d,
// RN(q + y * RN(a - q*d)) is the correctly rounded quotient (the residual a - q*d is exact in one FMA).  q0 = RN(a*y)
// is within 1.5 ulp, q1 within 0.5 ulp + 2^-53 ulp (hence faithful), q2 = RN(a/d) -- bit-identical to the `/` of the
// reference (src/deposition.jl:39-41, src/interpolation.jl:31-33) for every finite a, at 5 FP64 instructions instead of
// MUFU.RCP64H + 8 FP64 + range check + slow-path call (ncu, round 2: the cell-ordered particle kernels are bound by
// instruction issue, not by memory).  tests/test_exact_division.py checks the sequence against exact rational arithmetic.
__device__ __forceinline__ double div_exact(double a, double d, double rinv) {
    const double q0 = a * rinv;
    const double e0 = __fma_rn(-q0, d, a);
    const double q1 = __fma_rn(e0, rinv, q0);
    const double e1 = __fma_rn(-q1, d, a);
    return __fma_rn(e1, rinv, q1);
}
// the same steps in Float32 (y = RN_f32(1 / d) formed in Float32 on the host: Geom3::rinvf)
__device__ __forceinline__ float div_exact(float a, float d, float rinv) {
    const float q0 = a * rinv;
    const float e0 = __fmaf_rn(-q0, d, a);
    const float q1 = __fmaf_rn(e0, rinv, q0);
    const float e1 = __fmaf_rn(-q1, d, a);
    return __fmaf_rn(e1, rinv, q1);
}

// RN(1 / delta) in the working precision (a double-rounded Float32 reciprocal would not satisfy Markstein's condition)
template <typename W> __device__ __forceinline__ W geom_rinv(const Geom3& g, int a);
template <> __device__ __forceinline__ double geom_rinv<double>(const Geom3& g, int a) { return g.rinv[a]; }
template <> __device__ __forceinline__ float geom_rinv<float>(const Geom3& g, int a) { return g.rinvf[a]; }

__device__ __forceinline__ int floor_to_int(double t) { return __double2int_rd(t); }
__device__ __forceinline__ int floor_to_int(float t) { return __float2int_rd(t); }

// cell index (clamped to [0, n-2]) and fraction along one axis: floor, clamp and fraction as in the reference, with
// the floor taken by the float -> int conversion (saturating, so the integer clamp equals the clamp of the float value)
template <typename W>
__device__ __forceinline__ void locate_axis(W p, const Geom3& g, int a, int& i, W& f) {
    const W t = div_exact(p - (W)g.lo[a], (W)g.delta[a], geom_rinv<W>(g, a));
    i = min(max(floor_to_int(t), 0), g.n[a] - 2);
    f = t - (W)i;
}

template <typename W>
__device__ __forceinline__ void locate(W px, W py, W pz, const Geom3& g, CellW<W>& c) {
    locate_axis<W>(px, g, 0, c.i[0], c.f[0]);
    locate_axis<W>(py, g, 1, c.i[1], c.f[1]);
    locate_axis<W>(pz, g, 2, c.i[2], c.f[2]);
}

// z cell of a particle (same arithmetic as locate) and the slab filter of the gather kernels
template <typename W> __device__ __forceinline__ bool z_selected(W pz, const Geom3& g) {
    int iz;
    W f;
    locate_axis<W>(pz, g, 2, iz, f);
    return iz >= g.zlo && iz < g.zhi && !(iz >= g.exlo && iz < g.exhi);
}

// Result store shared by the gather kernels.  Plain interpolation writes the field value; the fused
// momentum kick updates the caller's array in place, p <- p + coef * E, with the
// product and the sum formed separately in W (no contraction: -fmad=false) and rounded to P once.
template <typename P, typename W>
__device__ __forceinline__ void put_result(P* __restrict__ arr, long long i, W val, const Kick& k, bool is_z) {
    if (k.on) val = (W)arr[i] + (W)(is_z ? k.cz : k.cxy) * val;
    st_stream(arr + i, (P)val);
}

}  // namespace scb
