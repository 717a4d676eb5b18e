// particles.cu -- the particle passes: cloud-in-cell deposit, trilinear gather, extrema, indices.
//
// replaces: src/deposition.jl:28-86 (deposit_particle!), :106-158 (deposit_kernel!/deposit_gpu!)
//           src/interpolation.jl:17-86 (interpolate_kernel!), src/mesh.jl:120-122 (extrema)
//
// Arithmetic follows the reference exactly: normalised
// coordinate with a true IEEE division in promote(P, T), floor, fraction, left-to-right weight
// products.  Compiled with -fmad=false (Julia does not contract a*b+c).  The only deviation is
// the clamp of the cell index to [0, n-2]: identical for in-range particles, and it removes the
// out-of-bounds write the reference performs for Float32 auto-bounds meshes.
#include "kernels.h"
#include "particle_common.cuh"

#include <cstdlib>

// minimum resident CTAs per SM requested for the Float64 lane-pair gather (register cap = 65536 / (256 * SCB_GATHER_MINB)).
// Measured at 1e8 particles / 256^3 with the run-time slab filter still in the kernel
// (profiles/r01_ab_gather_occupancy_s5.log): 4 CTAs 3.20 ms, 5 CTAs (48 registers, the compiler's own choice) 2.88 ms,
// 6 CTAs (40 registers, no spill) 2.85 ms, 8 CTAs (32 registers, 16 bytes spilled) 3.51 ms; with the filter compiled out
// (profiles/r01_ab_gather_filter_s5.log): 5 CTAs 2.61 ms, 6 CTAs 2.58 ms.
// build.py -D SCB_GATHER_MINB=<n> builds a variant for A/B timing
#ifndef SCB_GATHER_MINB
#define SCB_GATHER_MINB 6
#endif
#if SCB_GATHER_MINB > 0
#define SCB_GATHER_BOUNDS __launch_bounds__(256, SCB_GATHER_MINB)
#else
#define SCB_GATHER_BOUNDS __launch_bounds__(256)   // -D SCB_GATHER_MINB=0: the compiler's own register allocation
#endif

namespace scb {

// ---- deposit: one thread per particle, eight reductions into rho --------------------------
template <typename P, typename T>
__global__ void __launch_bounds__(256) k_deposit(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                  const P* __restrict__ z, const P* __restrict__ q,
                                                  T* __restrict__ rho, const Geom3 g) {
    using W = typename promote<P, T>::type;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        CellW<W> c;
        locate<W>((W)ld_stream(x + i), (W)ld_stream(y + i), (W)ld_stream(z + i), g, c);
        const W charge = (W)ld_stream(q + i);
        const W wx0 = (W)1 - c.f[0], wx1 = c.f[0];
        const W wy0 = (W)1 - c.f[1], wy1 = c.f[1];
        const W wz0 = (W)1 - c.f[2], wz1 = c.f[2];
        T* r = rho + c.i[0] + (long long)g.n[0] * (c.i[1] + (long long)g.n[1] * c.i[2]);
        const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
        // order of src/deposition.jl:67-74
        atomicAdd(r, (T)(charge * wx0 * wy0 * wz0));
        atomicAdd(r + 1, (T)(charge * wx1 * wy0 * wz0));
        atomicAdd(r + sy, (T)(charge * wx0 * wy1 * wz0));
        atomicAdd(r + sy + 1, (T)(charge * wx1 * wy1 * wz0));
        atomicAdd(r + sz, (T)(charge * wx0 * wy0 * wz1));
        atomicAdd(r + sz + 1, (T)(charge * wx1 * wy0 * wz1));
        atomicAdd(r + sz + sy, (T)(charge * wx0 * wy1 * wz1));
        atomicAdd(r + sz + sy + 1, (T)(charge * wx1 * wy1 * wz1));
    }
}

// ---- deposit, lane-pair variant ----------------------------------------------------------------
// ncu (round 1) shows the kernel above bound by the L2 reduction path (lts 84 %, dram 10 %): every
// reduction is its own 32-byte sector transaction.  A warp instruction's lanes that hit the same
// sector travel to L2 as ONE transaction, but the eight reductions of a particle are eight
// instructions.  Here two adjacent lanes share a particle (lane & 1 = x corner), so the two
// x-neighbours of every (y,z) corner pair go out in one reduction instruction and usually in one
// sector: 5 instead of 8 sector transactions per particle (measured: l1tex red sectors 5.0e8 for
// 1e8 particles).  Each lane forms its corners' values with the reference's expression
// ((q*wx)*wy)*wz, so the per-contribution values are bit-identical to k_deposit; only the (already
// unordered) accumulation order differs.  (An eight-lanes-per-particle version had the same
// sector count but spent a third of the LSU pipe on shuffles.)
template <typename P, typename T, bool ST>
__global__ void __launch_bounds__(256) k_deposit_pair(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                       const P* __restrict__ z, const P* __restrict__ q,
                                                       T* __restrict__ rho, const Geom3 g, const PLayout L) {
    using W = typename promote<P, T>::type;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const int kx = lane & 1;
    for (long long base = warp * 32; base < np; base += nwarps * 32) {
        const long long i = base + lane;
        W f0 = 0, f1 = 0, f2 = 0, charge = 0;
        long long off = 0;
        if (i < np) {
            CellW<W> c;
            locate<W>((W)ld_stream(x + pidx<ST>(i, L.x)), (W)ld_stream(y + pidx<ST>(i, L.y)),
                      (W)ld_stream(z + pidx<ST>(i, L.z)), g, c);
            charge = (W)ld_stream(q + pidx<ST>(i, L.q));
            f0 = c.f[0]; f1 = c.f[1]; f2 = c.f[2];
            off = c.i[0] + sy * c.i[1] + sz * c.i[2];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int src = 16 * h + (lane >> 1);
            const W fx = __shfl_sync(FULL, f0, src);
            const W fy = __shfl_sync(FULL, f1, src);
            const W fz = __shfl_sync(FULL, f2, src);
            const W qq = __shfl_sync(FULL, charge, src);
            const long long o = __shfl_sync(FULL, off, src);
            if (base + src < np) {
                const W one = (W)1;
                const W qx = qq * (kx ? fx : one - fx);        // charge * w_x
                const W qxy0 = qx * (one - fy), qxy1 = qx * fy;  // * w_y
                T* r = rho + o + kx;
                atomicAdd(r, (T)(qxy0 * (one - fz)));
                atomicAdd(r + sy, (T)(qxy1 * (one - fz)));
                atomicAdd(r + sz, (T)(qxy0 * fz));
                atomicAdd(r + sz + sy, (T)(qxy1 * fz));
            }
        }
    }
}

// ---- deposit, cell-tile variant ---------------------------------------------------------------
// The L2 reduction path charges per 32-byte sector transaction, not per value (k_deposit: 8 per
// particle, 4.1 ms; k_deposit_pair: 5 per particle, 2.65 ms at 1e8 particles / 256^3 Float64).  Here
// the particle does not update the grid nodes but a private accumulator laid out per CELL: tile
// (ix,iy,iz) = the four (x,y) corner values of cell (ix,iy) in plane iz, 4 values = one 32-byte
// sector in Float64.  Four adjacent lanes share a particle (lane&3 = x corner + 2*y corner) and issue
// two reduction instructions (plane iz, plane iz+1): 2 sector transactions per particle.  A second
// kernel folds the tiles into rho: node (i,j,k) = T(i,j,k)[0] + T(i-1,j,k)[1] + T(i,j-1,k)[2] +
// T(i-1,j-1,k)[3] (fixed order).  Per-contribution values are the reference's ((q*wx)*wy)*wz.
template <typename P, typename T, bool ST>
__global__ void __launch_bounds__(256) k_deposit_tiles(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                        const P* __restrict__ z, const P* __restrict__ q,
                                                        T* __restrict__ tiles, const Geom3 g, const PLayout L) {
    using W = typename promote<P, T>::type;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const int k = lane & 3, kx = k & 1, ky = k >> 1;
    const unsigned long long pol = l2_policy(g.l2_keep);
    // next group's particle data requested one iteration ahead (the reductions are fire-and-forget, so the loads of
    // x, y, z, q are the only latency on the critical path)
    P px = 0, py = 0, pz = 0, pq = 0;
    if (warp * 32 + lane < np) {
        const long long i0 = warp * 32 + lane;
        px = ld_stream(x + pidx<ST>(i0, L.x)); py = ld_stream(y + pidx<ST>(i0, L.y));
        pz = ld_stream(z + pidx<ST>(i0, L.z)); pq = ld_stream(q + pidx<ST>(i0, L.q));
    }
    for (long long base = warp * 32; base < np; base += nwarps * 32) {
        const long long i = base + lane;
        const long long inext = i + nwarps * 32;
        P nx_ = 0, ny_ = 0, nz_ = 0, nq_ = 0;
        if (inext < np) {
            nx_ = ld_stream(x + pidx<ST>(inext, L.x)); ny_ = ld_stream(y + pidx<ST>(inext, L.y));
            nz_ = ld_stream(z + pidx<ST>(inext, L.z)); nq_ = ld_stream(q + pidx<ST>(inext, L.q));
        }
        W f0 = 0, f1 = 0, f2 = 0, charge = 0;
        long long off = 0;
        if (i < np) {
            CellW<W> c;
            locate<W>((W)px, (W)py, (W)pz, g, c);
            charge = (W)pq;
            f0 = c.f[0]; f1 = c.f[1]; f2 = c.f[2];
            off = c.i[0] + sy * c.i[1] + sz * c.i[2];
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int src = 8 * h + (lane >> 2);
            const W fx = __shfl_sync(FULL, f0, src);
            const W fy = __shfl_sync(FULL, f1, src);
            const W fz = __shfl_sync(FULL, f2, src);
            const W qq = __shfl_sync(FULL, charge, src);
            const long long o = __shfl_sync(FULL, off, src);
            if (base + src < np) {
                const W one = (W)1;
                const W qxy = qq * (kx ? fx : one - fx) * (ky ? fy : one - fy);   // (charge * w_x) * w_y
                T* t = tiles + 4 * o + k;
                red_add_hint(t, (T)(qxy * (one - fz)), pol);
                red_add_hint(t + 4 * sz, (T)(qxy * fz), pol);
            }
        }
        px = nx_; py = ny_; pz = nz_; pq = nq_;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_fold_tiles(const T* __restrict__ tiles, T* __restrict__ rho, int nx, int ny,
                                                     long long ng, int accumulate) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ng) return;
    const int i = (int)(idx % nx), j = (int)((idx / nx) % ny);
    T s = __ldg(tiles + 4 * idx);
    if (i > 0) s += __ldg(tiles + 4 * (idx - 1) + 1);
    if (j > 0) s += __ldg(tiles + 4 * (idx - nx) + 2);
    if (i > 0 && j > 0) s += __ldg(tiles + 4 * (idx - nx - 1) + 3);
    rho[idx] = accumulate ? rho[idx] + s : s;
}

// ---- interpolate: one thread per particle, 24 gathers --------------------------------------
template <typename P, typename T, bool ST>
__global__ void __launch_bounds__(256) k_interpolate(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                      const P* __restrict__ z, const T* __restrict__ e,
                                                      const Geom3 g, P* __restrict__ ex, P* __restrict__ ey,
                                                      P* __restrict__ ez, const Kick kick, const PLayout L) {
    using W = typename promote<P, T>::type;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1], sc = sz * g.n[2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        CellW<W> c;
        locate<W>((W)ld_stream(x + pidx<ST>(i, L.x)), (W)ld_stream(y + pidx<ST>(i, L.y)),
                  (W)ld_stream(z + pidx<ST>(i, L.z)), g, c);
        const W dx = c.f[0], dy = c.f[1], dz = c.f[2];
        const W one = (W)1;
        // src/interpolation.jl:46-53
        const W w000 = (one - dx) * (one - dy) * (one - dz);
        const W w100 = dx * (one - dy) * (one - dz);
        const W w010 = (one - dx) * dy * (one - dz);
        const W w110 = dx * dy * (one - dz);
        const W w001 = (one - dx) * (one - dy) * dz;
        const W w101 = dx * (one - dy) * dz;
        const W w011 = (one - dx) * dy * dz;
        const W w111 = dx * dy * dz;
        const T* b = e + c.i[0] + sy * c.i[1] + sz * c.i[2];
        W out[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T* bk = b + k * sc;
            // src/interpolation.jl:56-85, left-to-right sum
            out[k] = (W)__ldg(bk) * w000 + (W)__ldg(bk + 1) * w100 + (W)__ldg(bk + sy) * w010 +
                     (W)__ldg(bk + sy + 1) * w110 + (W)__ldg(bk + sz) * w001 + (W)__ldg(bk + sz + 1) * w101 +
                     (W)__ldg(bk + sz + sy) * w011 + (W)__ldg(bk + sz + sy + 1) * w111;
        }
        put_result<P, W>(ex, pidx<ST>(i, L.ex), out[0], kick, false);
        put_result<P, W>(ey, pidx<ST>(i, L.ey), out[1], kick, false);
        put_result<P, W>(ez, pidx<ST>(i, L.ez), out[2], kick, true);
    }
}

// ---- interpolate, packed-field variant -------------------------------------------------------
// The 24 gathers per particle of the kernel above saturate the L1/L2 request path (ncu, round 1:
// l1tex 94 %, lts 76 %, dram 13 %).  The field is therefore first repacked node-major so that one
// 32-byte sector holds everything a particle needs from a node (Float64: {Ex,Ey,Ez,0}) or from an
// x-pair of nodes (Float32: {E(i), 0, E(i+1), 0}), and the gather becomes 8 (4) 256-bit loads.
// The arithmetic (weights, order of the eight products, left-to-right sum) is unchanged, so the
// results are bit-identical to k_interpolate.
__global__ void __launch_bounds__(256) k_pack_efield_f64(const double* __restrict__ e, double4* __restrict__ out,
                                                          long long ng, long long first, long long count) {
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + count) return;
    out[i] = make_double4(__ldg(e + i), __ldg(e + ng + i), __ldg(e + 2 * ng + i), 0.0);
}

__global__ void __launch_bounds__(256) k_pack_efield_f32(const float* __restrict__ e, float4* __restrict__ out,
                                                          long long ng, int nx, long long first, long long count) {
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + count) return;
    const bool last = (int)(i % nx) == nx - 1;
    const float4 a = make_float4(__ldg(e + i), __ldg(e + ng + i), __ldg(e + 2 * ng + i), 0.f);
    const float4 b = last ? make_float4(0.f, 0.f, 0.f, 0.f)
                          : make_float4(__ldg(e + i + 1), __ldg(e + ng + i + 1), __ldg(e + 2 * ng + i + 1), 0.f);
    out[2 * i] = a;
    out[2 * i + 1] = b;
}

__device__ __forceinline__ void ld256(const double4* p, double (&v)[4], unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ld256(const float4* p, float (&v)[8], unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p), "l"(pol));
}

template <typename P>
__global__ void __launch_bounds__(256) k_interpolate_packed_f64(long long np, const P* __restrict__ x,
                                                                 const P* __restrict__ y, const P* __restrict__ z,
                                                                 const double4* __restrict__ e, const Geom3 g,
                                                                 P* __restrict__ ex, P* __restrict__ ey,
                                                                 P* __restrict__ ez, const Kick kick) {
    using W = double;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const unsigned long long pol = l2_policy(g.l2_keep);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) {
        CellW<W> c;
        locate<W>((W)ld_stream(x + i), (W)ld_stream(y + i), (W)ld_stream(z + i), g, c);
        const double4* b = e + c.i[0] + sy * c.i[1] + sz * c.i[2];
        double n000[4], n100[4], n010[4], n110[4], n001[4], n101[4], n011[4], n111[4];
        ld256(b, n000, pol);
        ld256(b + 1, n100, pol);
        ld256(b + sy, n010, pol);
        ld256(b + sy + 1, n110, pol);
        ld256(b + sz, n001, pol);
        ld256(b + sz + 1, n101, pol);
        ld256(b + sz + sy, n011, pol);
        ld256(b + sz + sy + 1, n111, pol);
        const W dx = c.f[0], dy = c.f[1], dz = c.f[2], one = 1.0;
        const W w000 = (one - dx) * (one - dy) * (one - dz);
        const W w100 = dx * (one - dy) * (one - dz);
        const W w010 = (one - dx) * dy * (one - dz);
        const W w110 = dx * dy * (one - dz);
        const W w001 = (one - dx) * (one - dy) * dz;
        const W w101 = dx * (one - dy) * dz;
        const W w011 = (one - dx) * dy * dz;
        const W w111 = dx * dy * dz;
        W out[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            out[k] = n000[k] * w000 + n100[k] * w100 + n010[k] * w010 + n110[k] * w110 + n001[k] * w001 +
                     n101[k] * w101 + n011[k] * w011 + n111[k] * w111;
        put_result<P, W>(ex, i, out[0], kick, false);
        put_result<P, W>(ey, i, out[1], kick, false);
        put_result<P, W>(ez, i, out[2], kick, true);
    }
}

// Lane-pair variant for Float64 fields: ncu (round 1) shows the kernel above limited by L1 tag
// look-ups (one 128-byte line per lane per load, l1tex 92 %) before L2 (74 %).  Two adjacent lanes
// share a particle, lane&1 selects the x corner, so both x-neighbours of a (y,z) corner are
// fetched by ONE load instruction from (usually) one line.  Each lane forms the reference's
// weights (1-dx)*(1-dy)*(1-dz) ... and products; the eight products are summed as
// (p000+p010+p001+p011) + (p100+p110+p101+p111), i.e. in a different order than the reference's
// left-to-right sum (last-ulp differences, well inside the 1e-10 parity bar).
template <typename P>
__global__ void __launch_bounds__(256) k_interpolate_pair_f64(long long np, const P* __restrict__ x,
                                                               const P* __restrict__ y, const P* __restrict__ z,
                                                               const double4* __restrict__ e, const Geom3 g,
                                                               P* __restrict__ ex, P* __restrict__ ey,
                                                               P* __restrict__ ez, const Kick kick) {
    using W = double;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const unsigned long long pol = l2_policy(g.l2_keep);
    const int kx = lane & 1;
    for (long long base = warp * 32; base < np; base += nwarps * 32) {
        const long long i = base + lane;
        W f0 = 0, f1 = 0, f2 = 0;
        long long off = 0;
        if (i < np) {
            CellW<W> c;
            locate<W>((W)ld_stream(x + i), (W)ld_stream(y + i), (W)ld_stream(z + i), g, c);
            f0 = c.f[0]; f1 = c.f[1]; f2 = c.f[2];
            off = c.i[0] + sy * c.i[1] + sz * c.i[2];
        }
        W mine[3] = {0, 0, 0};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int src = 16 * h + (lane >> 1);
            const W dx = __shfl_sync(FULL, f0, src);
            const W dy = __shfl_sync(FULL, f1, src);
            const W dz = __shfl_sync(FULL, f2, src);
            const long long o = __shfl_sync(FULL, off, src);
            const double4* b = e + o + kx;
            double n00[4], n10[4], n01[4], n11[4];   // (y,z), (y+1,z), (y,z+1), (y+1,z+1) at this lane's x corner
            ld256(b, n00, pol);
            ld256(b + sy, n10, pol);
            ld256(b + sz, n01, pol);
            ld256(b + sz + sy, n11, pol);
            const W one = 1.0;
            const W wx = kx ? dx : one - dx;
            const W w00 = wx * (one - dy) * (one - dz);
            const W w10 = wx * dy * (one - dz);
            const W w01 = wx * (one - dy) * dz;
            const W w11 = wx * dy * dz;
            W acc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[k] = n00[k] * w00 + n10[k] * w10 + n01[k] * w01 + n11[k] * w11;
                const W other = __shfl_xor_sync(FULL, acc[k], 1);
                acc[k] = kx ? other + acc[k] : acc[k] + other;   // x0 part + x1 part on both lanes
                const W got = __shfl_sync(FULL, acc[k], 2 * (lane & 15));
                if ((lane >> 4) == h) mine[k] = got;
            }
        }
        if (i < np) {
            put_result<P, W>(ex, i, mine[0], kick, false);
            put_result<P, W>(ey, i, mine[1], kick, false);
            put_result<P, W>(ez, i, mine[2], kick, true);
        }
    }
}

// Lane pairs without broadcast / compaction shuffles.  The LSU data pipe is the busiest unit of the kernel above
// (ncu: 81 % of its wavefront rate) and shuffles travel through the same pipe: per 32 particles it spent 44
// wavefronts on shuffles (coordinates and offset broadcast to the pair, results gathered back for a coalesced
// store) next to ~144 on the gathers themselves.  Here both lanes of a pair load the particle's coordinates
// themselves (same address: one wavefront per 16 particles and array), locate it redundantly, and the even lane
// stores the result (16 active lanes, one 128-byte line): 12 shuffle wavefronts per 32 particles remain (the x0 + x1
// halves of the three components).  Same arithmetic, bit-identical results.
// FL = false (every caller except the slab passes of SCB_GATHER_OVERLAP) compiles the slab filter out: carrying it as a
// run-time flag cost the plain gather 10 % (2.62 -> 2.88 ms at 1e8 particles; z loaded first, x and y conditional).
template <typename P, bool ST, bool FL>
__global__ void SCB_GATHER_BOUNDS k_interpolate_pair2_f64(long long np, const P* __restrict__ x,
                                                                const P* __restrict__ y, const P* __restrict__ z,
                                                                const double4* __restrict__ e, const Geom3 g,
                                                                P* __restrict__ ex, P* __restrict__ ey,
                                                                P* __restrict__ ez, const Kick kick, const PLayout L) {
    using W = double;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const unsigned long long pol = l2_policy(g.l2_keep);
    const int kx = lane & 1;
    // the coordinates of the next group are requested before the gathers of the current one are consumed: the kernel is
    // latency-bound (ncu: long-scoreboard stalls 12 per issued instruction), and this takes the DRAM round trip of the
    // particle stream off the critical path of every iteration
    long long i = warp * 16 + (lane >> 1);
    const bool filt = FL && (g.zlo > 0 || g.zhi < g.n[2]);   // slab filter active (uniform)
    const bool zfirst = FL && g.zfirst != 0;   // x and y fetched on demand (slab passes that select few particles)
    P cx = 0, cy = 0, cz = 0;
    if (i < np) {
        cz = ld_stream(z + pidx<ST>(i, L.z));
        if (!zfirst) { cx = ld_stream(x + pidx<ST>(i, L.x)); cy = ld_stream(y + pidx<ST>(i, L.y)); }
    }
    for (; __any_sync(FULL, i < np); i += nwarps * 16) {
        const long long inext = i + nwarps * 16;
        P nx_ = 0, ny_ = 0, nz_ = 0;
        if (inext < np) {
            nz_ = ld_stream(z + pidx<ST>(inext, L.z));
            if (!zfirst) { nx_ = ld_stream(x + pidx<ST>(inext, L.x)); ny_ = ld_stream(y + pidx<ST>(inext, L.y)); }
        }
        const bool live = i < np && (!filt || z_selected<W>((W)cz, g));
        W acc[3] = {0, 0, 0};
        if (live) {
            if (zfirst) { cx = ld_stream(x + pidx<ST>(i, L.x)); cy = ld_stream(y + pidx<ST>(i, L.y)); }
            CellW<W> c;
            locate<W>((W)cx, (W)cy, (W)cz, g, c);
            const double4* b = e + (c.i[0] + sy * c.i[1] + sz * c.i[2]) + kx;
            double n00[4], n10[4], n01[4], n11[4];   // (y,z), (y+1,z), (y,z+1), (y+1,z+1) at this lane's x corner
            ld256(b, n00, pol);
            ld256(b + sy, n10, pol);
            ld256(b + sz, n01, pol);
            ld256(b + sz + sy, n11, pol);
            const W dx = c.f[0], dy = c.f[1], dz = c.f[2], one = 1.0;
            const W wx = kx ? dx : one - dx;
            const W w00 = wx * (one - dy) * (one - dz);
            const W w10 = wx * dy * (one - dz);
            const W w01 = wx * (one - dy) * dz;
            const W w11 = wx * dy * dz;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] = n00[k] * w00 + n10[k] * w10 + n01[k] * w01 + n11[k] * w11;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const W other = __shfl_xor_sync(FULL, acc[k], 1);
            acc[k] = kx ? other + acc[k] : acc[k] + other;   // x0 part + x1 part
        }
        if (live && kx == 0) {
            put_result<P, W>(ex, pidx<ST>(i, L.ex), acc[0], kick, false);
            put_result<P, W>(ey, pidx<ST>(i, L.ey), acc[1], kick, false);
            put_result<P, W>(ez, pidx<ST>(i, L.ez), acc[2], kick, true);
        }
        cx = nx_; cy = ny_; cz = nz_;
    }
}

template <typename P, bool ST, bool FL>
__global__ void __launch_bounds__(256) k_interpolate_packed_f32(long long np, const P* __restrict__ x,
                                                                 const P* __restrict__ y, const P* __restrict__ z,
                                                                 const float4* __restrict__ e, const Geom3 g,
                                                                 P* __restrict__ ex, P* __restrict__ ey,
                                                                 P* __restrict__ ez, const Kick kick, const PLayout L) {
    using W = typename promote<P, float>::type;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1];
    const unsigned long long pol = l2_policy(g.l2_keep);
    // coordinates requested one iteration ahead (latency-bound gather, see k_interpolate_pair2_f64)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool filt = FL && (g.zlo > 0 || g.zhi < g.n[2]);   // slab filter active (uniform)
    const bool zfirst = FL && g.zfirst != 0;   // x and y fetched on demand (slab passes that select few particles)
    P cx = 0, cy = 0, cz = 0;
    if (i < np) {
        cz = ld_stream(z + pidx<ST>(i, L.z));
        if (!zfirst) { cx = ld_stream(x + pidx<ST>(i, L.x)); cy = ld_stream(y + pidx<ST>(i, L.y)); }
    }
    for (; i < np; i += stride) {
        const long long inext = i + stride;
        P nx_ = 0, ny_ = 0, nz_ = 0;
        if (inext < np) {
            nz_ = ld_stream(z + pidx<ST>(inext, L.z));
            if (!zfirst) { nx_ = ld_stream(x + pidx<ST>(inext, L.x)); ny_ = ld_stream(y + pidx<ST>(inext, L.y)); }
        }
        if (filt && !z_selected<W>((W)cz, g)) {
            cx = nx_; cy = ny_; cz = nz_;
            continue;
        }
        if (zfirst) { cx = ld_stream(x + pidx<ST>(i, L.x)); cy = ld_stream(y + pidx<ST>(i, L.y)); }
        CellW<W> c;
        locate<W>((W)cx, (W)cy, (W)cz, g, c);
        const float4* b = e + 2 * (c.i[0] + sy * c.i[1] + sz * c.i[2]);
        float p00[8], p10[8], p01[8], p11[8];  // x-pairs at (y,z), (y+1,z), (y,z+1), (y+1,z+1)
        ld256(b, p00, pol);
        ld256(b + 2 * sy, p10, pol);
        ld256(b + 2 * sz, p01, pol);
        ld256(b + 2 * (sz + sy), p11, pol);
        const W dx = c.f[0], dy = c.f[1], dz = c.f[2], one = (W)1;
        const W w000 = (one - dx) * (one - dy) * (one - dz);
        const W w100 = dx * (one - dy) * (one - dz);
        const W w010 = (one - dx) * dy * (one - dz);
        const W w110 = dx * dy * (one - dz);
        const W w001 = (one - dx) * (one - dy) * dz;
        const W w101 = dx * (one - dy) * dz;
        const W w011 = (one - dx) * dy * dz;
        const W w111 = dx * dy * dz;
        W out[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            out[k] = (W)p00[k] * w000 + (W)p00[4 + k] * w100 + (W)p10[k] * w010 + (W)p10[4 + k] * w110 +
                     (W)p01[k] * w001 + (W)p01[4 + k] * w101 + (W)p11[k] * w011 + (W)p11[4 + k] * w111;
        put_result<P, W>(ex, pidx<ST>(i, L.ex), out[0], kick, false);
        put_result<P, W>(ey, pidx<ST>(i, L.ey), out[1], kick, false);
        put_result<P, W>(ez, pidx<ST>(i, L.ez), out[2], kick, true);
        cx = nx_; cy = ny_; cz = nz_;
    }
}

// ---- magnetic field of a bunch moving along +z (extension) -----------------
// B = (beta/c) z_hat x E:  Bx = -(beta/c) Ey,  By = (beta/c) Ex,  Bz = 0   (same SoA layout as efield)
template <typename T>
__global__ void __launch_bounds__(256) k_bfield(const T* __restrict__ e, T* __restrict__ b, long long ng, T boc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ng) return;
    const T ex = ld_stream(e + i), ey = ld_stream(e + ng + i);
    b[i] = -(boc * ey);
    b[ng + i] = boc * ex;
    b[2 * ng + i] = (T)0;
}

// ---- parity hook: unclamped cell indices ----------------------------------------------------
template <typename P, typename T>
__global__ void k_cell_index(long long np, const P* __restrict__ x, const P* __restrict__ y, const P* __restrict__ z,
                             const Geom3 g, long long* __restrict__ ix, long long* __restrict__ iy,
                             long long* __restrict__ iz) {
    using W = typename promote<P, T>::type;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    // the product kernels' quotient (div_exact), floor unclamped
    ix[i] = (long long)floor(div_exact((W)x[i] - (W)g.lo[0], (W)g.delta[0], geom_rinv<W>(g, 0)));
    iy[i] = (long long)floor(div_exact((W)y[i] - (W)g.lo[1], (W)g.delta[1], geom_rinv<W>(g, 1)));
    iz[i] = (long long)floor(div_exact((W)z[i] - (W)g.lo[2], (W)g.delta[2], geom_rinv<W>(g, 2)));
}

// ---- extrema -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_bounds_init(unsigned long long* out6) {
    if (threadIdx.x < 3) out6[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 6) out6[threadIdx.x] = 0ull;
}

template <typename P, bool ST>
__global__ void __launch_bounds__(256) k_bounds(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                 const P* __restrict__ z, unsigned long long* __restrict__ out6,
                                                 const PLayout L) {
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const long long stride = (long long)gridDim.x * blockDim.x;
    const P* arr[3] = {x, y, z};
    const long long es[3] = {L.x, L.y, L.z};
    // four independent streaming loads per array and iteration keep enough bytes in flight for HBM
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < np; i += 4 * stride) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v0 = (double)ld_stream(arr[a] + pidx<ST>(i, es[a])), v1 = (double)ld_stream(arr[a] + pidx<ST>(i + stride, es[a]));
            const double v2 = (double)ld_stream(arr[a] + pidx<ST>(i + 2 * stride, es[a])), v3 = (double)ld_stream(arr[a] + pidx<ST>(i + 3 * stride, es[a]));
            lo[a] = fmin(fmin(lo[a], fmin(v0, v1)), fmin(v2, v3));
            hi[a] = fmax(fmax(hi[a], fmax(v0, v1)), fmax(v2, v3));
        }
    }
    for (; i < np; i += stride) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = (double)ld_stream(arr[a] + pidx<ST>(i, es[a]));
            lo[a] = fmin(lo[a], v);
            hi[a] = fmax(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (lo[a] <= hi[a]) {
                atomicMin(out6 + a, order_key(lo[a]));
                atomicMax(out6 + 3 + a, order_key(hi[a]));
            }
        }
    }
}

// ---- L2 probes (measurement only, scb_debug_l2_probe) -------------------------------------------
// Ceilings for the scattered 32-byte sector traffic that bounds the particle passes: the gather issues 256-bit
// no-allocate loads of random node records, the deposit four-lane fp64 reductions into random tile sectors.  The probes
// issue exactly those instructions at pseudo-random sector addresses of a buffer (L2-resident when it is small) with
// no other work, eight independent operations per thread and iteration; bench.py reports the particle kernels'
// achieved sector rates against these measured ceilings next to the HBM-byte roofline.
__device__ __forceinline__ unsigned long long lcg(unsigned long long s) {
    return s * 6364136223846793005ull + 1442695040888963407ull;
}

__global__ void __launch_bounds__(256) k_probe_sector_reads(const double4* __restrict__ buf, unsigned long long nrec,
                                                             int iters, double* __restrict__ sink) {
    // two adjacent lanes fetch two adjacent records (the x-neighbours of the lane-pair gather): 32 sectors and about 20
    // distinct 128-byte lines per warp instruction, like k_interpolate_pair2_f64
    unsigned long long s = ((((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1) + 1) * 0x9E3779B97F4A7C15ull;
    const unsigned long long pol = l2_policy(1);
    const int kx = threadIdx.x & 1;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double v[8][4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = lcg(s);
            ld256(buf + __umulhi((unsigned)(s >> 32), (unsigned)(nrec - 1)) + kx, v[k], pol);   // uniform in [0, nrec-1)
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k][0];
    }
    if (acc == 0.123456789) sink[0] = acc;   // never true for a zeroed buffer; keeps the loads observable
}

__global__ void __launch_bounds__(256) k_probe_sector_reds(double* __restrict__ buf, unsigned long long ntile, int iters) {
    // four adjacent lanes share one random tile (one 32-byte sector), like k_deposit_tiles
    unsigned long long s = ((((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2) + 1) * 0x9E3779B97F4A7C15ull;
    const unsigned long long pol = l2_policy(1);
    const int k = threadIdx.x & 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = lcg(s);
            red_add_hint(buf + 4ull * __umulhi((unsigned)(s >> 32), (unsigned)ntile) + k, 1.0, pol);
        }
    }
}

cudaError_t launch_probe(int mode, void* buf, size_t bytes, int iters, unsigned grid, cudaStream_t s, double* ops) {
    const unsigned long long nrec = bytes / 32;
    if (mode == 0) {
        k_probe_sector_reads<<<grid, 256, 0, s>>>((const double4*)buf, nrec, iters, (double*)buf);
        *ops = (double)grid * 256.0 * iters * 8.0;          // one sector (one record) per lane and load
    } else {
        k_probe_sector_reds<<<grid, 256, 0, s>>>((double*)buf, nrec, iters);
        *ops = (double)grid * 64.0 * iters * 8.0;           // one sector per four-lane group and reduction
    }
    return cudaGetLastError();
}

// ---- launchers -----------------------------------------------------------------------------
static inline unsigned particle_grid(long long np, int bs, int per_sm) {
    long long want = (np + bs - 1) / bs;
    long long cap = 148LL * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

// CTAs per SM of the grid-stride tile deposit and packed gather (tuning: SCB_DEPOSIT_PER_SM / SCB_GATHER_PER_SM).
// Measured at 1e8 particles / 256^3 Float64 (profiles/r01_ab_particle_grid_s5.log): a grid of exactly the resident
// CTAs runs all warps in lockstep through the load / gather / store phases (gather 4.2 ms at 5 per SM, deposit 1.97 ms at
// 4 per SM); from 64 per SM on the curve is flat (gather 2.85 / 2.83 / 2.825 / 2.84 ms at 64 / 128 / 256 / 512).
static int per_sm_env(const char* name) {
    const char* e = getenv(name);
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 256;
}

#define SCB_DISPATCH_PT(CALL)                                                  \
    if (pdt == 0 && mdt == 0) { CALL(float, float) }                           \
    else if (pdt == 0 && mdt == 1) { CALL(float, double) }                     \
    else if (pdt == 1 && mdt == 0) { CALL(double, float) }                     \
    else { CALL(double, double) }

// ST = true instantiations serve the scb_*_strided entry points; contiguous callers keep the ST = false kernels
#define SCB_LAYOUT(...)                                                        \
    if (lay) { const PLayout L = *lay; constexpr bool ST = true; __VA_ARGS__ } \
    else { const PLayout L{}; constexpr bool ST = false; __VA_ARGS__ }

// FL = true instantiations carry the z-slab filter of the overlapped field gather (needs `const Geom3& g` in scope)
#define SCB_FILTER(...)                                                        \
    if (g.zlo > 0 || g.zhi < g.n[2] || g.zfirst != 0) { constexpr bool FL = true; __VA_ARGS__ } \
    else { constexpr bool FL = false; __VA_ARGS__ }

cudaError_t launch_deposit(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                           const void* q, void* rho, const Geom3& g, int mode, cudaStream_t s, const PLayout* lay) {
    if (np <= 0) return cudaSuccess;
    const unsigned grid = particle_grid(np, 256, 64);
    if (mode == 1 && !lay) {
#define CALL(P, T) k_deposit<P, T><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g);
        SCB_DISPATCH_PT(CALL)
#undef CALL
    } else {
#define CALL(P, T) SCB_LAYOUT(k_deposit_pair<P, T, ST><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g, L);)
        SCB_DISPATCH_PT(CALL)
#undef CALL
    }
    return cudaGetLastError();
}

cudaError_t launch_interpolate(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                               const void* efield, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                               const Kick& kick, const PLayout* lay) {
    if (np <= 0) return cudaSuccess;
    const unsigned grid = particle_grid(np, 256, 64);
#define CALL(P, T) SCB_LAYOUT(k_interpolate<P, T, ST><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const T*)efield, g, (P*)ex, (P*)ey, (P*)ez, kick, L);)
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_deposit_tiles(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                 const void* q, void* tiles, void* rho, const Geom3& g, int accumulate, cudaStream_t s,
                                 const PLayout* lay) {
    const long long ng = (long long)g.n[0] * g.n[1] * g.n[2];
    cudaError_t e = cudaMemsetAsync(tiles, 0, (size_t)4 * ng * (mdt == 1 ? 8 : 4), s);
    if (e != cudaSuccess) return e;
    if (np > 0) {
        static const int per_sm = per_sm_env("SCB_DEPOSIT_PER_SM");
        const unsigned grid = particle_grid(np, 256, per_sm);
#define CALL(P, T) SCB_LAYOUT(k_deposit_tiles<P, T, ST><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)tiles, g, L);)
        SCB_DISPATCH_PT(CALL)
#undef CALL
    }
    const unsigned fgrid = (unsigned)((ng + 255) / 256);
    if (mdt == 1) k_fold_tiles<double><<<fgrid, 256, 0, s>>>((const double*)tiles, (double*)rho, g.n[0], g.n[1], ng, accumulate);
    else k_fold_tiles<float><<<fgrid, 256, 0, s>>>((const float*)tiles, (float*)rho, g.n[0], g.n[1], ng, accumulate);
    return cudaGetLastError();
}

int interp_mode() {
    static const int m = [] { const char* e = getenv("SCB_INTERP_MODE"); return e ? atoi(e) : 0; }();
    return m;
}

size_t packed_bytes_per_node(int) { return 32; }

cudaError_t launch_pack_efield(int mdt, const void* efield, void* packed, const Geom3& g, cudaStream_t s, long long first,
                               long long count) {
    const long long ng = (long long)g.n[0] * g.n[1] * g.n[2];
    if (count < 0) count = ng - first;
    if (count <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (mdt == 1) k_pack_efield_f64<<<grid, 256, 0, s>>>((const double*)efield, (double4*)packed, ng, first, count);
    else k_pack_efield_f32<<<grid, 256, 0, s>>>((const float*)efield, (float4*)packed, ng, g.n[0], first, count);
    return cudaGetLastError();
}

cudaError_t launch_interpolate_packed(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                      const void* packed, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                                      const Kick& kick, const PLayout* lay) {
    if (np <= 0) return cudaSuccess;
    static const int per_sm = per_sm_env("SCB_GATHER_PER_SM");
    const unsigned grid = particle_grid(np, 256, per_sm);
    static const int imode = interp_mode();
    const bool thread_per_particle = imode == 1 && !lay;   // the tuning variants exist for contiguous arrays only
    if (mdt == 1 && thread_per_particle) {
        if (pdt == 1) k_interpolate_packed_f64<double><<<grid, 256, 0, s>>>(np, (const double*)x, (const double*)y, (const double*)z, (const double4*)packed, g, (double*)ex, (double*)ey, (double*)ez, kick);
        else k_interpolate_packed_f64<float><<<grid, 256, 0, s>>>(np, (const float*)x, (const float*)y, (const float*)z, (const double4*)packed, g, (float*)ex, (float*)ey, (float*)ez, kick);
    } else if (mdt == 1 && (imode != 3 || lay)) {
        if (pdt == 1) { SCB_LAYOUT(SCB_FILTER(k_interpolate_pair2_f64<double, ST, FL><<<grid, 256, 0, s>>>(np, (const double*)x, (const double*)y, (const double*)z, (const double4*)packed, g, (double*)ex, (double*)ey, (double*)ez, kick, L);)) }
        else { SCB_LAYOUT(SCB_FILTER(k_interpolate_pair2_f64<float, ST, FL><<<grid, 256, 0, s>>>(np, (const float*)x, (const float*)y, (const float*)z, (const double4*)packed, g, (float*)ex, (float*)ey, (float*)ez, kick, L);)) }
    } else if (mdt == 1) {
        if (pdt == 1) k_interpolate_pair_f64<double><<<grid, 256, 0, s>>>(np, (const double*)x, (const double*)y, (const double*)z, (const double4*)packed, g, (double*)ex, (double*)ey, (double*)ez, kick);
        else k_interpolate_pair_f64<float><<<grid, 256, 0, s>>>(np, (const float*)x, (const float*)y, (const float*)z, (const double4*)packed, g, (float*)ex, (float*)ey, (float*)ez, kick);
    } else {
        if (pdt == 1) { SCB_LAYOUT(SCB_FILTER(k_interpolate_packed_f32<double, ST, FL><<<grid, 256, 0, s>>>(np, (const double*)x, (const double*)y, (const double*)z, (const float4*)packed, g, (double*)ex, (double*)ey, (double*)ez, kick, L);)) }
        else { SCB_LAYOUT(SCB_FILTER(k_interpolate_packed_f32<float, ST, FL><<<grid, 256, 0, s>>>(np, (const float*)x, (const float*)y, (const float*)z, (const float4*)packed, g, (float*)ex, (float*)ey, (float*)ez, kick, L);)) }
    }
    return cudaGetLastError();
}

cudaError_t launch_bfield(int mdt, const void* efield, void* bfield, long long ng, double beta_over_c, cudaStream_t s) {
    const unsigned grid = (unsigned)((ng + 255) / 256);
    if (mdt == 1) k_bfield<double><<<grid, 256, 0, s>>>((const double*)efield, (double*)bfield, ng, beta_over_c);
    else k_bfield<float><<<grid, 256, 0, s>>>((const float*)efield, (float*)bfield, ng, (float)beta_over_c);
    return cudaGetLastError();
}

cudaError_t launch_cell_index(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                              const Geom3& g, long long* ix, long long* iy, long long* iz, cudaStream_t s) {
    if (np <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((np + 255) / 256);
#define CALL(P, T) k_cell_index<P, T><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, g, ix, iy, iz);
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_bounds(int pdt, long long np, const void* x, const void* y, const void* z, double* out6,
                          cudaStream_t s, const PLayout* lay) {
    unsigned long long* o = reinterpret_cast<unsigned long long*>(out6);
    k_bounds_init<<<1, 32, 0, s>>>(o);
    if (np > 0) {
        const unsigned grid = particle_grid(np, 256, 16);
        if (pdt == 0) { SCB_LAYOUT(k_bounds<float, ST><<<grid, 256, 0, s>>>(np, (const float*)x, (const float*)y, (const float*)z, o, L);) }
        else { SCB_LAYOUT(k_bounds<double, ST><<<grid, 256, 0, s>>>(np, (const double*)x, (const double*)y, (const double*)z, o, L);) }
    }
    return cudaGetLastError();
}

}  // namespace scb
