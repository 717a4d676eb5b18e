// sorted.cu -- the cell-ordered regime of the particle passes.
//
// replaces (for bunches the caller keeps ordered by cell): src/deposition.jl:28-86, 106-158 (deposit) and
//           src/interpolation.jl:17-86 (gather); scb_sort_particles has no counterpart in the reference.
//
// With particles in random order every corner update / corner read is its own 32-byte L2 transaction and both particle
// passes sit at the ceiling of that machinery.  A tracking loop can keep its bunch ordered by cell
// (particles move a fraction of a cell per step, so the order degrades slowly and a re-sort every K steps is enough).
// This file holds what that regime needs:
//   * k_cell_keys + an LSD radix sort (k_radix_hist / k_scan_* / k_radix_scatter) -> the permutation that orders a
//     bunch by linear cell index, and k_permute that applies it to any number of per-particle arrays;
//   * k_deposit_runs: every lane walks 8 CONSECUTIVE particles, accumulates the eight corner values of the current
//     cell in registers and only leaves the registers when the cell changes; the lanes of a warp then combine their
//     open runs with a segmented shuffle scan (warp-aggregated reduction) and one lane per distinct cell issues the
//     eight reductions.  Correct for ANY order (a cell change just ends a run); fast when neighbours share cells;
//   * k_interpolate_runs: the same walk for the gather -- the 24 field values of a cell stay in registers while
//     consecutive particles remain in it, coordinates and results move as 256-bit vectors.
// The per-particle arithmetic (locate, weights, products, order of the eight terms) is the reference's; compiled with
// -fmad=false like particles.cu.
#include "kernels.h"
#include "particle_common.cuh"

#include <cstdint>
#include <cstdlib>

namespace scb {

namespace {
constexpr unsigned FULL = 0xffffffffu;
#ifndef SCB_SORT_THREADS
#define SCB_SORT_THREADS 512
#endif
constexpr int SORT_THREADS = SCB_SORT_THREADS;
#ifndef SCB_SORT_KPT
#define SCB_SORT_KPT 8
#endif
constexpr int SORT_KPT = SCB_SORT_KPT;                // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_KPT;    // keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_IPT;   // 4096 counters per CTA
}  // namespace

// ---- cell keys ----------------------------------------------------------------------------------------------------
// key = linear index of the particle's cell, ix + nx*(iy + ny*iz), from the same locate() as the deposit and the gather
// (clamped to [0, n-2] per axis), so "sorted by key" is exactly "consecutive particles share cells" for those kernels.
// One CTA per sort tile: besides the keys it leaves the tile's digit histogram of the FIRST radix pass (hist != nullptr),
// which saves that pass its own read of the keys.
template <typename P, typename T, bool ST>
__global__ void __launch_bounds__(SORT_THREADS) k_cell_keys(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                             const P* __restrict__ z, const Geom3 g, unsigned* __restrict__ keys,
                                                             const PLayout L, unsigned mask, unsigned* __restrict__ hist,
                                                             int ntiles) {
    using W = typename promote<P, T>::type;
    __shared__ unsigned sh[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 256) sh[tid] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_TILE + warp * (SORT_KPT * 32) + lane;
#pragma unroll 2
    for (int r = 0; r < SORT_KPT; ++r) {
        const long long i = base + r * 32;
        unsigned d = 0xffffffffu;
        if (i < np) {
            CellW<W> c;
            locate<W>((W)ld_stream(x + pidx<ST>(i, L.x)), (W)ld_stream(y + pidx<ST>(i, L.y)),
                      (W)ld_stream(z + pidx<ST>(i, L.z)), g, c);
            const unsigned key = (unsigned)c.i[0] + (unsigned)g.n[0] * ((unsigned)c.i[1] + (unsigned)g.n[1] * (unsigned)c.i[2]);
            keys[i] = key;
            d = key & mask;
        }
        if (hist) {
            const unsigned d0 = __shfl_sync(FULL, d, 0);
            if (__all_sync(FULL, d == d0)) {
                if (lane == 0 && d0 != 0xffffffffu) atomicAdd(&sh[d0], 32u);
            } else if (d != 0xffffffffu) {
                atomicAdd(&sh[d], 1u);
            }
        }
    }
    __syncthreads();
    if (hist && (unsigned)tid <= mask) hist[(size_t)tid * ntiles + blockIdx.x] = sh[tid];
}

// ---- LSD radix sort of (key, index) pairs, 32-bit keys, up to 8 bits per pass -------------------------------------
// Three kernels per pass: per-tile digit histograms (digit-major, so that ONE exclusive scan over the whole table
// yields every tile's global base per digit), the scan, and the scatter.  Ranks inside a tile come from
// __match_any_sync on the digit (cost independent of the digit distribution: a bunch that is already almost sorted --
// the common case in a tracking loop -- has one digit value per tile in the last pass).  A warp owns 32 * SORT_KPT consecutive
// keys of the tile and walks them in SORT_KPT rounds of 32, so (warp, round, lane) order = index order and the sort is stable.
__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const unsigned* __restrict__ keys, long long n, int shift,
                                                              unsigned mask, unsigned* __restrict__ hist, int ntiles) {
    __shared__ unsigned sh[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 256) sh[tid] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_TILE + warp * (SORT_KPT * 32) + lane;
    // all of the thread's keys first (independent loads), then the counting: one vote per round catches the case that bounds shared
    // atomics (every lane on the same counter -- the last pass of an almost ordered bunch), everything else goes through
    // native integer atomics (the match-based count of the first version made this kernel latency-bound: 0.60 ms for
    // 0.4 GB)
    unsigned d[SORT_KPT];
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        const long long i = base + r * 32;
        d[r] = i < n ? ((__ldg(keys + i) >> shift) & mask) : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        const unsigned d0 = __shfl_sync(FULL, d[r], 0);
        if (__all_sync(FULL, d[r] == d0)) {
            if (lane == 0 && d0 != 0xffffffffu) atomicAdd(&sh[d0], 32u);
        } else if (d[r] != 0xffffffffu) {
            atomicAdd(&sh[d[r]], 1u);
        }
    }
    __syncthreads();
    if ((unsigned)tid <= mask) hist[(size_t)tid * ntiles + blockIdx.x] = sh[tid];
}

// exclusive scan of `n` counters in place: per-chunk sums, scan of the sums by one CTA, per-chunk scan + base
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total, unsigned* s_warp /*[32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < nw ? s_warp[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const unsigned wbase = warp > 0 ? s_warp[warp - 1] : 0u;
    if (total) *total = s_warp[nw - 1];
    const unsigned r = wbase + inc - v;
    __syncthreads();   // s_warp may be reused by the caller
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const unsigned* __restrict__ a, long long n,
                                                               unsigned* __restrict__ partial) {
    __shared__ unsigned s_warp[32];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_IPT;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k)
        if (base + k < n) s += a[base + k];
    unsigned total;
    (void)block_exclusive_scan(s, &total, s_warp);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_top(unsigned* __restrict__ partial, int nparts) {
    __shared__ unsigned s_warp[32];
    unsigned carry = 0;
    for (int base = 0; base < nparts; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nparts ? partial[i] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive_scan(v, &total, s_warp);
        if (i < nparts) partial[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(unsigned* __restrict__ a, long long n,
                                                              const unsigned* __restrict__ partial) {
    __shared__ unsigned s_warp[32];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_IPT;
    unsigned v[SCAN_IPT];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        v[k] = base + k < n ? a[base + k] : 0u;
        s += v[k];
    }
    unsigned run = partial[blockIdx.x] + block_exclusive_scan(s, nullptr, s_warp);
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < n) a[base + k] = run;
        run += v[k];
    }
}

// vals_in == nullptr: the values are the key indices themselves (first pass); keys_out == nullptr: the ordered keys are
// not needed (last pass)
__global__ void __launch_bounds__(SORT_THREADS, 1024 / SORT_THREADS) k_radix_scatter(const unsigned* __restrict__ keys_in,
                                                                    const unsigned* __restrict__ vals_in,
                                                                    unsigned* __restrict__ keys_out,
                                                                    unsigned* __restrict__ vals_out, long long n, int shift,
                                                                    unsigned mask, const unsigned* __restrict__ offs,
                                                                    int ntiles) {
    extern __shared__ unsigned smem[];
    unsigned* s_keys = smem;                          // SORT_TILE
    unsigned* s_vals = smem + SORT_TILE;              // SORT_TILE
    unsigned* whist = smem + 2 * SORT_TILE;           // SORT_WARPS * 256: per-warp digit counts, then per-warp bases
    unsigned* dstart = whist + SORT_WARPS * 256;      // 256: first slot of the digit inside the ordered tile
    unsigned* gbase = dstart + 256;                   // 256: global position of slot i of digit d = gbase[d] + i
    __shared__ unsigned s_warp[32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile0 = (long long)blockIdx.x * SORT_TILE;
    const int count = (int)((n - tile0) < SORT_TILE ? (n - tile0) : SORT_TILE);
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) whist[i] = 0;
    const int first = warp * (SORT_KPT * 32) + lane;   // tile-local index of this thread's round-0 key
    unsigned k[SORT_KPT];
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        const int li = first + r * 32;
        // keys beyond the end take the largest digit: they rank behind every real key of the tile
        k[r] = li < count ? __ldg(keys_in + tile0 + li) : 0xffffffffu;
    }
    __syncthreads();
    unsigned* wh = whist + warp * 256;
    // rank of every key among the warp's keys with the same digit, in (round, lane) order.  The 16 match operations are
    // independent and issue back to back; the per-digit running counts then advance with one returning shared-memory
    // atomic per round and distinct digit (issued by the first lane of each group of equal digits), and the old values
    // reach the other lanes of the group with one shuffle per round.
    unsigned info[SORT_KPT];   // leader | rank inside the round << 5 | group size << 10
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        const unsigned peers = __match_any_sync(FULL, (k[r] >> shift) & mask);
        info[r] = (unsigned)(__ffs(peers) - 1) | ((unsigned)__popc(peers & ((1u << lane) - 1u)) << 5) |
                  ((unsigned)__popc(peers) << 10);
    }
    unsigned rk[SORT_KPT];
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        rk[r] = 0;
        if ((info[r] & 31u) == (unsigned)lane) rk[r] = atomicAdd(&wh[(k[r] >> shift) & mask], info[r] >> 10);
        __syncwarp();   // orders this round's atomics before the next round's (ranks must follow the round order)
    }
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) rk[r] = __shfl_sync(FULL, rk[r], info[r] & 31u) + ((info[r] >> 5) & 31u);
    __syncthreads();
    // per digit: exclusive prefix over the warps, total of the tile
    unsigned total = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const unsigned t = whist[w * 256 + tid];
            whist[w * 256 + tid] = total;
            total += t;
        }
    }
    const unsigned ex = block_exclusive_scan(total, nullptr, s_warp);   // threads >= 256 contribute 0 behind the digits
    if (tid < 256) {
        dstart[tid] = ex;
        gbase[tid] = ((unsigned)tid <= mask ? offs[(size_t)tid * ntiles + blockIdx.x] : 0u) - ex;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_KPT; ++r) {
        const int li = first + r * 32;
        const unsigned d = (k[r] >> shift) & mask;
        const unsigned pos = dstart[d] + wh[d] + rk[r];
        s_keys[pos] = k[r];
        s_vals[pos] = li < count ? (vals_in ? __ldg(vals_in + tile0 + li) : (unsigned)(tile0 + li)) : 0u;
    }
    __syncthreads();
    for (int i = tid; i < count; i += SORT_THREADS) {
        const unsigned key = s_keys[i];
        const unsigned d = (key >> shift) & mask;
        const unsigned pos = gbase[d] + (unsigned)i;
        if (keys_out) keys_out[pos] = key;
        vals_out[pos] = s_vals[i];
    }
}

// dst_f[i] = src_f[perm[i]] for up to SCB_MAX_PERMUTE_FIELDS arrays of one element size (4 or 8 bytes)
struct PermuteArgs {
    const void* src[8];
    void* dst[8];
};

template <typename V>
__global__ void __launch_bounds__(256) k_permute(long long n, const unsigned* __restrict__ perm, int nf, const PermuteArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned p = ld_stream(perm + i);
        V v[8];
#pragma unroll
        for (int f = 0; f < 8; ++f)
            if (f < nf) v[f] = __ldg(static_cast<const V*>(a.src[f]) + p);
#pragma unroll
        for (int f = 0; f < 8; ++f)
            if (f < nf) st_stream(static_cast<V*>(a.dst[f]) + i, v[f]);
    }
}

// fraction of sampled neighbour pairs (i, i+1) that lie in the same or the x-adjacent cell: ~1 for a cell-ordered bunch,
// ~0 for a random one.  out[0] += hits, out[1] += pairs (unsigned long long counters, zeroed by the launcher).
template <typename P, typename T>
__global__ void __launch_bounds__(256) k_order_probe(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                      const P* __restrict__ z, const Geom3 g, long long block_stride,
                                                      unsigned long long* __restrict__ out) {
    using W = typename promote<P, T>::type;
    const long long i = (long long)blockIdx.x * block_stride + threadIdx.x;
    int key = -2;
    if (i < np) {
        CellW<W> c;
        locate<W>((W)x[i], (W)y[i], (W)z[i], g, c);
        key = c.i[0] + g.n[0] * (c.i[1] + g.n[1] * c.i[2]);
    }
    const int nxt = __shfl_down_sync(FULL, key, 1);
    const bool pair = (threadIdx.x & 31) != 31 && key >= 0 && nxt >= 0;
    const bool hit = pair && (nxt - key >= -1 && nxt - key <= 1);
    const unsigned hits = __popc(__ballot_sync(FULL, hit)), pairs = __popc(__ballot_sync(FULL, pair));
    if ((threadIdx.x & 31) == 0 && pairs) {
        atomicAdd(out, (unsigned long long)hits);
        atomicAdd(out + 1, (unsigned long long)pairs);
    }
}

// ---- 256-bit particle loads / stores ---------------------------------------------------------------------------------
// eight consecutive particles of one array: two v4.f64 or one v8.f32 (pointer 32-byte aligned)
__device__ __forceinline__ void ld8(const double* p, double (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]) : "l"(p + 4));
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, const double (&v)[2]) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void st4(float* p, const float (&v)[2]) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// ---- deposit over runs of particles that share a cell -----------------------------------------------------------------
// Where a finished run goes.  GlobalSink: eight reductions into rho.  WindowSink: a shared-memory tile of rho nodes that
// the CTA keeps around the cells its particles are in (k_deposit_window) -- runs inside the tile are added there with
// shared-memory atomics, anything else falls through to the global reductions, so results never depend on the tile.
template <typename T, typename W>
struct GlobalSink {
    T* __restrict__ rho;
    long long sy, sz;
    unsigned long long pol;
    __device__ __forceinline__ int key(const int (&i)[3]) const { return i[0] + (int)sy * i[1] + (int)sz * i[2]; }
    __device__ __forceinline__ void global_add(int cell, const W (&s)[8]) const {
        T* r = rho + cell;
        // corner order of src/deposition.jl:67-74
        red_add_hint(r, (T)s[0], pol);
        red_add_hint(r + 1, (T)s[1], pol);
        red_add_hint(r + sy, (T)s[2], pol);
        red_add_hint(r + sy + 1, (T)s[3], pol);
        red_add_hint(r + sz, (T)s[4], pol);
        red_add_hint(r + sz + 1, (T)s[5], pol);
        red_add_hint(r + sz + sy, (T)s[6], pol);
        red_add_hint(r + sz + sy + 1, (T)s[7], pol);
    }
    __device__ __forceinline__ void flush(int k, const W (&s)[8]) const { global_add(k, s); }
};

template <typename T, typename W>
struct WindowSink {
    GlobalSink<T, W> g;
    W* win;                 // [wz][wy][wx] nodes, origin (x0, y0, z0)
    int x0, y0, z0, wx, wy, wz;
    // cell coordinates packed into one word (grid dimensions are at most 1024): equal keys <=> equal cells
    __device__ __forceinline__ int key(const int (&i)[3]) const { return i[0] | (i[1] << 10) | (i[2] << 20); }
    __device__ __forceinline__ void flush(int k, const W (&s)[8]) const {
        const int ix = k & 1023, iy = (k >> 10) & 1023, iz = k >> 20;
        const int lx = ix - x0, ly = iy - y0, lz = iz - z0;
        if ((unsigned)lx < (unsigned)(wx - 1) && (unsigned)ly < (unsigned)(wy - 1) && (unsigned)lz < (unsigned)(wz - 1)) {
            W* w = win + lx + wx * (ly + wy * lz);
            const int py = wx, pz = wx * wy;
            atomicAdd(w, s[0]);
            atomicAdd(w + 1, s[1]);
            atomicAdd(w + py, s[2]);
            atomicAdd(w + py + 1, s[3]);
            atomicAdd(w + pz, s[4]);
            atomicAdd(w + pz + 1, s[5]);
            atomicAdd(w + pz + py, s[6]);
            atomicAdd(w + pz + py + 1, s[7]);
        } else {
            g.global_add(ix + (int)g.sy * iy + (int)g.sz * iz, s);
        }
    }
};

// one lane's walk over its (up to) 8 consecutive particles: the corner sums of the current cell stay in s[], a cell
// change sends them to the sink.  ALL: every one of the 8 exists (no per-particle bound checks).
// A particle that differs from the current run while the NEXT particle is back in it is an interloper (a neighbour
// that drifted into another cell since the last sort): it goes to the sink on its own and the run stays open -- one
// flush instead of two (measured after a drift of 0.1 cell: see profiles/r02_sorted_regime_*.json).
template <typename P, typename W, bool ALL, typename Sink>
__device__ __forceinline__ void walk_deposit(const P (&px)[8], const P (&py)[8], const P (&pz)[8], const P (&pq)[8], int cnt,
                                             const Geom3& g, const Sink& sink, int& cur, W (&s)[8]) {
    CellW<W> c, cn;
    locate<W>((W)px[0], (W)py[0], (W)pz[0], g, cn);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (ALL || j < cnt) {
            c = cn;
            const int cell = sink.key(c.i);
            int cell_next = -2;
            if (j + 1 < 8 && (ALL || j + 1 < cnt)) {
                locate<W>((W)px[j + 1 < 8 ? j + 1 : 7], (W)py[j + 1 < 8 ? j + 1 : 7], (W)pz[j + 1 < 8 ? j + 1 : 7], g, cn);
                cell_next = sink.key(cn.i);
            }
            const W one = (W)1, charge = (W)pq[j];
            const W qx0 = charge * (one - c.f[0]), qx1 = charge * c.f[0];              // charge * w_x
            const W wy0 = one - c.f[1], wy1 = c.f[1], wz0 = one - c.f[2], wz1 = c.f[2];
            const W qxy00 = qx0 * wy0, qxy10 = qx1 * wy0, qxy01 = qx0 * wy1, qxy11 = qx1 * wy1;   // * w_y
            const W v[8] = {qxy00 * wz0, qxy10 * wz0, qxy01 * wz0, qxy11 * wz0,
                            qxy00 * wz1, qxy10 * wz1, qxy01 * wz1, qxy11 * wz1};       // ((q*wx)*wy)*wz
            const bool same = cell == cur;
            if (!same && cur >= 0 && cell_next == cur) {
                sink.flush(cell, v);          // interloper: the run continues with the next particle
            } else {
                if (!same && cur >= 0) sink.flush(cur, s);
                cur = cell;
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] = (same ? s[k] : (W)0) + v[k];
            }
        }
    }
}

// the lanes' open runs: adjacent lanes with the same cell form a segment; inclusive segmented scan, the last lane of
// every segment holds its sum and hands it to the sink (warp-aggregated reduction)
template <typename W, typename Sink>
__device__ __forceinline__ void combine_open_runs(int lane, int cur, W (&s)[8], const Sink& sink) {
    const int prev = __shfl_up_sync(FULL, cur, 1), next = __shfl_down_sync(FULL, cur, 1);
    const unsigned heads = __ballot_sync(FULL, lane == 0 || prev != cur);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const bool take = lane - o >= start;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const W t = __shfl_up_sync(FULL, s[k], o);
            s[k] += take ? t : (W)0;
        }
    }
    if (cur >= 0 && (lane == 31 || next != cur)) sink.flush(cur, s);
}

template <typename P, bool VEC>
__device__ __forceinline__ void load_lane8(const P* __restrict__ x, const P* __restrict__ y, const P* __restrict__ z,
                                           const P* __restrict__ q, long long i0, int cnt, bool full, P (&px)[8], P (&py)[8],
                                           P (&pz)[8], P (&pq)[8]) {
    if (VEC && full) {
        ld8(x + i0, px);
        ld8(y + i0, py);
        ld8(z + i0, pz);
        ld8(q + i0, pq);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = j < cnt;
            px[j] = ok ? ld_stream(x + i0 + j) : (P)0;
            py[j] = ok ? ld_stream(y + i0 + j) : (P)0;
            pz[j] = ok ? ld_stream(z + i0 + j) : (P)0;
            pq[j] = ok ? ld_stream(q + i0 + j) : (P)0;
        }
    }
}

// VEC: all four arrays are 32-byte aligned (256-bit loads); otherwise element loads.
template <typename P, typename T, bool VEC>
__global__ void __launch_bounds__(256, 2) k_deposit_runs(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                          const P* __restrict__ z, const P* __restrict__ q,
                                                          T* __restrict__ rho, const Geom3 g) {
    using W = typename promote<P, T>::type;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const GlobalSink<T, W> sink{rho, g.n[0], (long long)g.n[0] * g.n[1], l2_policy(g.l2_keep)};
    for (long long wbase = warp * 256; wbase < np; wbase += nwarps * 256) {
        const long long i0 = wbase + lane * 8;
        const bool full = wbase + 256 <= np;   // warp-uniform
        const int cnt = (int)(np - i0 >= 8 ? 8 : (np - i0 > 0 ? np - i0 : 0));
        P px[8], py[8], pz[8], pq[8];
        load_lane8<P, VEC>(x, y, z, q, i0, cnt, full, px, py, pz, pq);
        int cur = -1;
        W s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = (W)0;
        if (full) walk_deposit<P, W, true>(px, py, pz, pq, 8, g, sink, cur, s);
        else walk_deposit<P, W, false>(px, py, pz, pq, cnt, g, sink, cur, s);
        combine_open_runs<W>(lane, cur, s, sink);
    }
}

// ---- the same walk with a shared-memory tile under it -------------------------------------------------------------------
// A CTA takes DW_ITERS * 2048 CONSECUTIVE particles and keeps a tile of rho nodes (DW_WX x DW_WY x DW_WZ, 16 KB in
// Float64) centred on the cell of the first particle of the current 2048: one row of cells of margin in y and z, a
// few cells behind and many ahead in x (the order is x-fastest).  Finished runs inside the tile are added with
// shared-memory atomics; the tile is flushed to rho with coalesced reductions when it has to move and at the end.
// For a freshly sorted bunch this changes little (runs are long, few reach memory at all).  It matters once the bunch
// has DRIFTED since its last sort: a particle that moved to a neighbouring cell interrupts its neighbours' run, and
// each interruption costs two flushes -- 16 global reductions with k_deposit_runs (measured at 1e8 particles / 256^3:
// 0.66 ms freshly sorted, 2.4 ms after a drift of 0.1 cell, 4.1 ms after 0.3 cell), a few shared-memory atomics here,
// because the neighbouring cells are in the tile.  Particles outside the tile (sparse regions, where 2048 consecutive
// particles span many rows) take the global path as before.
constexpr int DW_ITERS = 8, DW_WX = 128, DW_WY = 4, DW_WZ = 4, DW_AHEAD = 40, DW_BEHIND = 8;

template <typename P, typename T, bool VEC>
__global__ void __launch_bounds__(256, 2) k_deposit_window(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                            const P* __restrict__ z, const P* __restrict__ q,
                                                            T* __restrict__ rho, const Geom3 g) {
    using W = typename promote<P, T>::type;
    __shared__ W win[DW_WX * DW_WY * DW_WZ];
    __shared__ int s_new[2][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WindowSink<T, W> sink;
    sink.g = GlobalSink<T, W>{rho, g.n[0], (long long)g.n[0] * g.n[1], l2_policy(g.l2_keep)};
    sink.win = win;
    sink.wx = g.n[0] < DW_WX ? g.n[0] : DW_WX;
    sink.wy = g.n[1] < DW_WY ? g.n[1] : DW_WY;
    sink.wz = g.n[2] < DW_WZ ? g.n[2] : DW_WZ;
    sink.x0 = sink.y0 = sink.z0 = -1;   // no tile yet
    const int wn = sink.wx * sink.wy * sink.wz;
    for (int i = tid; i < wn; i += 256) win[i] = (W)0;
    auto flush_tile = [&]() {   // all threads; tile -> rho (coalesced along x), tile zeroed
        if (sink.x0 < 0) return;
        for (int i = tid; i < wn; i += 256) {
            const W v = win[i];
            if (v != (W)0) {
                const int lx = i % sink.wx, r = i / sink.wx, ly = r % sink.wy, lz = r / sink.wy;
                red_add_hint(rho + (sink.x0 + lx) + sink.g.sy * (sink.y0 + ly) + sink.g.sz * (sink.z0 + lz), (T)v, sink.g.pol);
                win[i] = (W)0;
            }
        }
    };
    const long long cta_base = (long long)blockIdx.x * (2048LL * DW_ITERS);
    for (int it = 0; it < DW_ITERS && cta_base + it * 2048LL < np; ++it) {
        const long long wbase = cta_base + it * 2048LL + warp * 256;
        const long long i0 = wbase + lane * 8;
        const bool full = wbase + 256 <= np;   // warp-uniform
        const int cnt = (int)(np - i0 >= 8 ? 8 : (np - i0 > 0 ? np - i0 : 0));
        P px[8], py[8], pz[8], pq[8];
        load_lane8<P, VEC>(x, y, z, q, i0, cnt, full, px, py, pz, pq);
        // does the tile still sit around the first particle of these 2048?  (decided by thread 0, which holds it)
        int need = 0;
        if (tid == 0) {
            CellW<W> c;
            locate<W>((W)px[0], (W)py[0], (W)pz[0], g, c);
            const int nx0 = min(max(c.i[0] - DW_BEHIND, 0), g.n[0] - sink.wx);
            const int ny0 = min(max(c.i[1] - 1, 0), g.n[1] - sink.wy);
            const int nz0 = min(max(c.i[2] - 1, 0), g.n[2] - sink.wz);
            const int lx = c.i[0] - sink.x0;
            const bool x_ok = sink.x0 >= 0 && (lx >= 2 || sink.x0 == 0) &&
                              (lx <= sink.wx - 2 - DW_AHEAD || sink.x0 + sink.wx == g.n[0]) && lx >= 0 && lx <= sink.wx - 2;
            need = !(x_ok && ny0 == sink.y0 && nz0 == sink.z0);
            if (need) {
                s_new[it & 1][0] = nx0;
                s_new[it & 1][1] = ny0;
                s_new[it & 1][2] = nz0;
            }
        }
        // (the barrier also orders the previous iteration's shared-memory atomics before a flush of the tile)
        if (__syncthreads_or(need)) {
            flush_tile();
            __syncthreads();
            sink.x0 = s_new[it & 1][0];
            sink.y0 = s_new[it & 1][1];
            sink.z0 = s_new[it & 1][2];
        }
        int cur = -1;
        W s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = (W)0;
        if (full) walk_deposit<P, W, true>(px, py, pz, pq, 8, g, sink, cur, s);
        else walk_deposit<P, W, false>(px, py, pz, pq, cnt, g, sink, cur, s);
        combine_open_runs<W>(lane, cur, s, sink);
    }
    __syncthreads();
    flush_tile();
}

// ---- gather over runs of particles that share a cell ------------------------------------------------------------------
// Every lane walks 4 consecutive particles; the 24 field values of the current cell stay in registers while the cell
// does not change.  Weights, products and the left-to-right sum are those of src/interpolation.jl:46-85 (bit-identical
// to k_interpolate).  Results leave as one 256-bit (Float64) / 128-bit (Float32) store per component.
// tuning (build.py -D ...): particles per lane (4 or 2), CTA size, minimum resident CTAs, results stored one by one
#ifndef SCB_GR_M
#define SCB_GR_M 4
#endif
// measured at 1e8 particles / 256^3 Float64, cell-ordered (profiles/r02_ab_gather_runs.log, first table: no prefetch):
// 256 threads x 2 CTAs (128 registers, 428 bytes spilled) 1.38 ms; 128 threads x 3 CTAs (168 registers) 1.21 ms;
// 2 particles per lane 1.77 ms; element stores as results are formed 2.05 ms (partial-sector writes); one thread per
// particle from L1 1.60 ms.  Second table, with the coordinate prefetch: 128 x 3 (168 registers, 108 bytes spilled)
// 1.31 ms; 128 x 2 (252 registers, no spill) 1.15 ms; 256 x 1 1.20 ms; prefetch compiled out 1.25 ms
#ifndef SCB_GR_THREADS
#define SCB_GR_THREADS 128
#endif
#ifndef SCB_GR_MINB
#define SCB_GR_MINB 2
#endif
#ifndef SCB_GR_STORE_NOW
#define SCB_GR_STORE_NOW 0
#endif
#ifndef SCB_GR_PREFETCH
#define SCB_GR_PREFETCH 1
#endif
constexpr int GR_M = SCB_GR_M;

// ALL: every one of the lane's GR_M particles exists.  NOW: results are stored as they are formed (element stores)
// instead of being kept for one vector store per component.
template <typename P, typename T, typename W, bool ALL, bool NOW>
__device__ __forceinline__ void walk_gather(const P (&px)[GR_M], const P (&py)[GR_M], const P (&pz)[GR_M], int cnt,
                                            const Geom3& g, const T* __restrict__ e, long long sy, long long sz,
                                            long long sc, P (&ox)[GR_M], P (&oy)[GR_M], P (&oz)[GR_M], const Kick& kick,
                                            P* __restrict__ ex, P* __restrict__ ey, P* __restrict__ ez) {
    int cur = -1;
    T n[3][8];
#pragma unroll
    for (int j = 0; j < GR_M; ++j) {
        if (!NOW) ox[j] = oy[j] = oz[j] = (P)0;
        if (ALL || j < cnt) {
            CellW<W> c;
            locate<W>((W)px[j], (W)py[j], (W)pz[j], g, c);
            const int cell = c.i[0] + (int)sy * c.i[1] + (int)sz * c.i[2];
            if (cell != cur) {
                cur = cell;
                const T* b = e + cell;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const T* bk = b + k * sc;
                    n[k][0] = __ldg(bk);
                    n[k][1] = __ldg(bk + 1);
                    n[k][2] = __ldg(bk + sy);
                    n[k][3] = __ldg(bk + sy + 1);
                    n[k][4] = __ldg(bk + sz);
                    n[k][5] = __ldg(bk + sz + 1);
                    n[k][6] = __ldg(bk + sz + sy);
                    n[k][7] = __ldg(bk + sz + sy + 1);
                }
            }
            const W dx = c.f[0], dy = c.f[1], dz = c.f[2], one = (W)1;
            // src/interpolation.jl:46-53
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
            for (int k = 0; k < 3; ++k)   // src/interpolation.jl:56-85, left-to-right sum
                out[k] = (W)n[k][0] * w000 + (W)n[k][1] * w100 + (W)n[k][2] * w010 + (W)n[k][3] * w110 +
                         (W)n[k][4] * w001 + (W)n[k][5] * w101 + (W)n[k][6] * w011 + (W)n[k][7] * w111;
            if (NOW) {
                put_result<P, W>(ex, j, out[0], kick, false);
                put_result<P, W>(ey, j, out[1], kick, false);
                put_result<P, W>(ez, j, out[2], kick, true);
            } else {
                ox[j] = (P)out[0];
                oy[j] = (P)out[1];
                oz[j] = (P)out[2];
            }
        }
    }
}

template <typename P, typename T, bool VEC>
__global__ void __launch_bounds__(SCB_GR_THREADS, SCB_GR_MINB) k_interpolate_runs(long long np, const P* __restrict__ x, const P* __restrict__ y,
                                                              const P* __restrict__ z, const T* __restrict__ e,
                                                              const Geom3 g, P* __restrict__ ex, P* __restrict__ ey,
                                                              P* __restrict__ ez, const Kick kick) {
    using W = typename promote<P, T>::type;
    constexpr int PER_WARP = 32 * GR_M;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sy = g.n[0], sz = (long long)g.n[0] * g.n[1], sc = sz * g.n[2];
    // coordinates of the warp's next group are requested before the current group is worked on: ncu on the version
    // without it showed 4.2 long-scoreboard stalls per issued instruction at 12 warps per SM (every iteration exposed
    // the DRAM round trip of its three coordinate loads)
    auto load_group = [&](long long wb, P (&a)[GR_M], P (&b)[GR_M], P (&c)[GR_M]) {
        const long long j0 = wb + lane * GR_M;
        if (VEC && wb + PER_WARP <= np) {
            ld4(x + j0, a);
            ld4(y + j0, b);
            ld4(z + j0, c);
        } else {
            const int m = (int)(np - j0 >= GR_M ? GR_M : (np - j0 > 0 ? np - j0 : 0));
#pragma unroll
            for (int j = 0; j < GR_M; ++j) {
                const bool ok = j < m;
                a[j] = ok ? ld_stream(x + j0 + j) : (P)0;
                b[j] = ok ? ld_stream(y + j0 + j) : (P)0;
                c[j] = ok ? ld_stream(z + j0 + j) : (P)0;
            }
        }
    };
    P px[GR_M], py[GR_M], pz[GR_M];
    long long wbase = warp * PER_WARP;
    if (wbase < np) load_group(wbase, px, py, pz);
    for (; wbase < np; wbase += nwarps * PER_WARP) {
        const long long i0 = wbase + lane * GR_M;
        const bool full = wbase + PER_WARP <= np;   // warp-uniform
        const int cnt = (int)(np - i0 >= GR_M ? GR_M : (np - i0 > 0 ? np - i0 : 0));
        const long long wnext = wbase + nwarps * PER_WARP;
        P qx[GR_M], qy[GR_M], qz[GR_M];
#pragma unroll
        for (int j = 0; j < GR_M; ++j) qx[j] = qy[j] = qz[j] = (P)0;
        if (SCB_GR_PREFETCH && wnext < np) load_group(wnext, qx, qy, qz);
        P ox[GR_M], oy[GR_M], oz[GR_M];
        constexpr bool NOW = SCB_GR_STORE_NOW != 0;
        if (VEC && full && !kick.on) {
            walk_gather<P, T, W, true, NOW>(px, py, pz, GR_M, g, e, sy, sz, sc, ox, oy, oz, kick, ex + i0, ey + i0, ez + i0);
            if (!NOW) {
                st4(ex + i0, ox);
                st4(ey + i0, oy);
                st4(ez + i0, oz);
            }
        } else {
            // tail of the bunch, unaligned arrays, fused momentum kick (p <- p + coef * E): element stores
            walk_gather<P, T, W, false, true>(px, py, pz, cnt, g, e, sy, sz, sc, ox, oy, oz, kick, ex + i0, ey + i0, ez + i0);
        }
        if (SCB_GR_PREFETCH) {
#pragma unroll
            for (int j = 0; j < GR_M; ++j) { px[j] = qx[j]; py[j] = qy[j]; pz[j] = qz[j]; }
        } else if (wnext < np) {
            load_group(wnext, px, py, pz);
        }
    }
}

// ---- launchers ------------------------------------------------------------------------------------------------------
#define SCB_DISPATCH_PT(CALL)                                                  \
    if (pdt == 0 && mdt == 0) { CALL(float, float) }                           \
    else if (pdt == 0 && mdt == 1) { CALL(float, double) }                     \
    else if (pdt == 1 && mdt == 0) { CALL(double, float) }                     \
    else { CALL(double, double) }

static inline unsigned capped_grid(long long items, int per_block, int per_sm) {
    long long want = (items + per_block - 1) / per_block;
    const long long cap = 148LL * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

static inline bool aligned32(const void* a, const void* b, const void* c, const void* d) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d)) & 31u) == 0;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static inline int sort_digit_bits(int key_bits) {
    if (key_bits < 1) key_bits = 1;
    const int passes = (key_bits + 7) / 8;
    return (key_bits + passes - 1) / passes;
}

// scratch: the sort scratch of launch_sort_pairs (sort_scratch_bytes); fills its first key buffer and the histogram table
// of the first pass
cudaError_t launch_cell_keys(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const Geom3& g,
                             void* scratch, int key_bits, cudaStream_t s, const PLayout* lay) {
    if (np <= 0) return cudaSuccess;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const int ntiles = (int)((np + SORT_TILE - 1) / SORT_TILE);
    unsigned* keys = static_cast<unsigned*>(scratch);
    unsigned* hist = reinterpret_cast<unsigned*>(static_cast<char*>(scratch) + 3 * up((size_t)np * 4));
    const unsigned mask = (1u << sort_digit_bits(key_bits)) - 1u;
#define CALL(P, T)                                                                                                        \
    if (lay) k_cell_keys<P, T, true><<<ntiles, SORT_THREADS, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, g, keys, *lay, mask, hist, ntiles); \
    else k_cell_keys<P, T, false><<<ntiles, SORT_THREADS, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, g, keys, PLayout{}, mask, hist, ntiles);
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

size_t sort_scratch_bytes(long long n) {
    const long long ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    const long long nh = 256 * ntiles;
    const long long nparts = (nh + SCAN_CHUNK - 1) / SCAN_CHUNK;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    // two key buffers, one value buffer, the histogram table, the scan partials
    return 3 * up((size_t)n * 4) + up((size_t)nh * 4) + up((size_t)nparts * 4);
}

// orders the indices 0..n-1 by key (stable); the first key buffer and the first pass's histogram table inside `scratch`
// come from launch_cell_keys.  perm_out receives the permutation.  `key_bits` significant bits.  *launches: kernel count.
cudaError_t launch_sort_pairs(void* scratch, long long n, int key_bits, unsigned* perm_out, cudaStream_t s, int* launches) {
    if (launches) *launches = 0;
    if (n <= 0) return cudaSuccess;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const int ntiles = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const long long nh_max = 256LL * ntiles;
    char* base = static_cast<char*>(scratch);
    unsigned* keys[2] = {reinterpret_cast<unsigned*>(base), reinterpret_cast<unsigned*>(base + up((size_t)n * 4))};
    unsigned* vals_scratch = reinterpret_cast<unsigned*>(base + 2 * up((size_t)n * 4));
    unsigned* hist = reinterpret_cast<unsigned*>(base + 3 * up((size_t)n * 4));
    unsigned* partial = reinterpret_cast<unsigned*>(base + 3 * up((size_t)n * 4) + up((size_t)nh_max * 4));
    if (key_bits < 1) key_bits = 1;
    const int passes = (key_bits + 7) / 8;
    const int bits = sort_digit_bits(key_bits);
    const size_t smem = (size_t)(2 * SORT_TILE + SORT_WARPS * 256 + 512) * 4;
    {   // per call: the attribute belongs to the current device, and a process may drive several
        cudaError_t e = cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // value buffers alternate so that the LAST pass writes perm_out
    unsigned* vbuf[2];
    vbuf[(passes - 1) & 1] = perm_out;
    vbuf[passes & 1] = vals_scratch;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * bits;
        const unsigned mask = (1u << bits) - 1u;
        const long long nh = (long long)(mask + 1) * ntiles;
        const int nparts = (int)((nh + SCAN_CHUNK - 1) / SCAN_CHUNK);
        const unsigned* kin = keys[p & 1];
        unsigned* kout = p == passes - 1 ? nullptr : keys[(p + 1) & 1];
        const unsigned* vin = p == 0 ? nullptr : vbuf[(p - 1) & 1];
        unsigned* vout = vbuf[p & 1];
        if (p > 0) k_radix_hist<<<ntiles, SORT_THREADS, 0, s>>>(kin, n, shift, mask, hist, ntiles);
        k_scan_reduce<<<nparts, SCAN_THREADS, 0, s>>>(hist, nh, partial);
        k_scan_top<<<1, 1024, 0, s>>>(partial, nparts);
        k_scan_apply<<<nparts, SCAN_THREADS, 0, s>>>(hist, nh, partial);
        k_radix_scatter<<<ntiles, SORT_THREADS, smem, s>>>(kin, vin, kout, vout, n, shift, mask, hist, ntiles);
        if (launches) *launches += p > 0 ? 5 : 4;
    }
    return cudaGetLastError();
}

cudaError_t launch_permute(int elem_bytes, long long n, const unsigned* perm, int nf, const void* const* src, void* const* dst,
                           cudaStream_t s) {
    if (n <= 0 || nf <= 0) return cudaSuccess;
    PermuteArgs a{};
    for (int f = 0; f < nf && f < 8; ++f) {
        a.src[f] = src[f];
        a.dst[f] = dst[f];
    }
    const unsigned grid = capped_grid(n, 256, 64);
    if (elem_bytes == 8) k_permute<double><<<grid, 256, 0, s>>>(n, perm, nf, a);
    else k_permute<float><<<grid, 256, 0, s>>>(n, perm, nf, a);
    return cudaGetLastError();
}

cudaError_t launch_order_probe(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const Geom3& g,
                               unsigned long long* counters2, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(counters2, 0, 2 * sizeof(unsigned long long), s);
    if (e != cudaSuccess || np < 2) return e;
    // at most 4096 sampled windows of 256 consecutive particles, evenly spread over the bunch
    const long long nblocks_all = (np + 255) / 256;
    const long long nb = nblocks_all < 4096 ? nblocks_all : 4096;
    const long long block_stride = (nblocks_all / nb) * 256;
#define CALL(P, T) k_order_probe<P, T><<<(unsigned)nb, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, g, block_stride, counters2);
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_deposit_runs(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const void* q,
                                void* rho, const Geom3& g, cudaStream_t s, bool tile) {
    if (np <= 0) return cudaSuccess;
    static const int per_sm = env_int("SCB_RUNS_PER_SM", 64);
    const unsigned grid = capped_grid(np, 256 * 8, per_sm);
    const bool vec = aligned32(x, y, z, q);
    // shared-memory tile under the walk: measured slower (shared Float64 atomics are compare-and-swap loops), opt-in
    const unsigned wgrid = (unsigned)((np + 2048LL * DW_ITERS - 1) / (2048LL * DW_ITERS));
#define CALL(P, T)                                                                                                         \
    if (tile) {                                                                                                            \
        if (vec) k_deposit_window<P, T, true><<<wgrid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g); \
        else k_deposit_window<P, T, false><<<wgrid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g); \
    } else {                                                                                                               \
        if (vec) k_deposit_runs<P, T, true><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g); \
        else k_deposit_runs<P, T, false><<<grid, 256, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const P*)q, (T*)rho, g); \
    }
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_interpolate_runs(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                    const void* efield, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                                    const Kick& kick) {
    if (np <= 0) return cudaSuccess;
    static const int per_sm = env_int("SCB_RUNS_PER_SM", 64);
    const unsigned grid = capped_grid(np, SCB_GR_THREADS * GR_M, per_sm * 256 / SCB_GR_THREADS);
    const bool vec = aligned32(x, y, z, ex) && aligned32(ey, ez, ey, ez);
#define CALL(P, T)                                                                                                         \
    if (vec) k_interpolate_runs<P, T, true><<<grid, SCB_GR_THREADS, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const T*)efield, g, (P*)ex, (P*)ey, (P*)ez, kick); \
    else k_interpolate_runs<P, T, false><<<grid, SCB_GR_THREADS, 0, s>>>(np, (const P*)x, (const P*)y, (const P*)z, (const T*)efield, g, (P*)ex, (P*)ey, (P*)ez, kick);
    SCB_DISPATCH_PT(CALL)
#undef CALL
    return cudaGetLastError();
}

}  // namespace scb
