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

template <typename W>
__device__ __forceinline__ void locate(W px, W py, W pz, const Geom3& g, CellW<W>& c) {
    const W p[3] = {px, py, pz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const W t = (p[a] - (W)g.lo[a]) / (W)g.delta[a];
        W fl = floor(t);
        fl = fmin(fmax(fl, (W)0), (W)(g.n[a] - 2));
        c.i[a] = (int)fl;
        c.f[a] = t - fl;
    }
}

// z cell of a particle (same arithmetic as locate) and the slab filter of the gather kernels
template <typename W> __device__ __forceinline__ bool z_selected(W pz, const Geom3& g) {
    const W t = (pz - (W)g.lo[2]) / (W)g.delta[2];
    const int iz = (int)fmin(fmax(floor(t), (W)0), (W)(g.n[2] - 2));
    return iz >= g.zlo && iz < g.zhi && !(iz >= g.exlo && iz < g.exhi);
}

}  // namespace scb
