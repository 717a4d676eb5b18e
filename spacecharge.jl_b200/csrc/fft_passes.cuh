// fft_passes.cuh -- the five pass groups of the pruned FFT convolution and the
// generic passes used to build the Green-function spectrum.
//
// Data layout of every complex intermediate: x (kx) fastest with pitch PX (a multiple of 8,
// >= Lx/2+1), then y, then z.  Only kx in [0, Lx/2] is kept (real input => Hermitian spectrum).
//
//   F1  x_r2c      rho (nx,ny,nz) real  -> A [kx][y ][z ]   zero-pad nx->Lx fused into the load,
//                                                          two real lines per complex transform
//   F2  lines<-1>  A                    -> B [kx][ky][z ]   zero-pad ny->Ly fused into the load
//   Z   z_fused    B                    -> C_c[kx][ky][z ]  pad nz->Lz, forward, x Green_c
//                                                          (+ mirrored spectrum x image Green_c),
//                                                          inverse, keep first nz; c = 0,1,2
//   B2  lines<+1>  C_c                  -> D_c[kx][y ][z ]  inverse along y, keep first ny
//   B3  x_c2r      D_c                  -> E_c (nx,ny,nz)   inverse along x, keep first nx,
//                                                          x FPEI/(Lx Ly Lz) fused into the store
//
// replaces: src/solvers/free_space.jl:68-99 (fill!, embed, 7 in-place C2C FFTs, multiply,
// scale, extract) of the reference.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is resolved at run time in api.cu)
#include "fft_engine.cuh"

namespace scb {

// threads that advance in lockstep along the contiguous axis in the strided passes
#ifndef SCB_LTX_BIG
#define SCB_LTX_BIG 8
#endif
__host__ __device__ constexpr int tx_for(int N) {
    return N >= 2048 ? 2 : N == 1024 ? 4 : N >= 256 ? SCB_LTX_BIG : N == 128 ? 16 : 32;
}
// lockstep lines of the fused z pass (its register footprint is twice that of a plain pass, so
// it runs narrower CTAs to keep two of them resident per SM)
#ifndef SCB_ZTX_BIG
#define SCB_ZTX_BIG 4
#endif
#ifndef SCB_Z_MINBLOCKS
#define SCB_Z_MINBLOCKS 2
#endif
#ifndef SCB_Z_MINBLOCKS_F32
#define SCB_Z_MINBLOCKS_F32 3
#endif
template <typename T> __host__ __device__ constexpr int z_minblocks() { return sizeof(T) == 4 ? SCB_Z_MINBLOCKS_F32 : SCB_Z_MINBLOCKS; }
__host__ __device__ constexpr int tz_for(int N) {
    return N >= 2048 ? 2 : N == 1024 ? 4 : N >= 256 ? SCB_ZTX_BIG : N == 128 ? 16 : 32;
}
// line pairs per CTA in the x passes
__host__ __device__ constexpr int lp_for(int N) { return (N / 8) >= 256 ? 1 : 256 / (N / 8); }

#ifndef SCB_MAX_RANKS
#define SCB_MAX_RANKS 16
#endif

enum GreenKind : int { GREEN_FREE = 0, GREEN_CATHODE = 1, GREEN_FULL = 2 };

// ------------------------------------------------------------------------------------------
// generic strided complex pass
template <typename T>
struct LinesParams {
    const cx_t<T>* in;
    cx_t<T>* out;
    const cx_t<T>* tw;
    int n_in, n_out;    // positions >= n_in read as zero; bins >= n_out are not stored
    int ninner;         // valid elements along the contiguous axis
    long long in_sline, in_souter, out_sline, out_souter;
    long long in_scomp, out_scomp;  // blockIdx.z selects the field component
    // slab-decomposed (multi-GPU) runs: the line is cut into blocks of `split` positions, one per
    // peer rank; position pos lives at (pos % split)*sline + (pos / split)*sblock.  0 = contiguous.
    int in_split, out_split;
    long long in_sblock, out_sblock;
    // peer-memory variant of the output split: block b is stored through out_peer[b] (a pointer into
    // rank b's receive buffer, mapped over NVLink) -- the all-to-all is fused into the pass
    // use_peers == 2: the peer is chosen by the OUTER index instead (kx-slab solve, B2: plane z goes to the owner of its
    // z slab): outer o is stored through out_peer[o / out_osplit] at outer index o % out_osplit
    int use_peers;
    int out_osplit;
    cx_t<T>* out_peer[SCB_MAX_RANKS];
    // input known to be even (fold_sign = +1) or odd (-1) about index 0 (mod N): only positions
    // 0..N/2 are stored, position pos > N/2 is read as fold_sign * in[N - pos]
    int in_fold;
    T fold_sign;
    T scale;
    // Green-spectrum build on real-symmetric data: the complex line is a PAIR of real lines (a + i b).  Both are
    // even => the output is (A, B) with A, B real.  Both odd => A = i A~, B = i B~ and the transform is -B~ + i A~;
    // out_rot multiplies it by -i so that (A~, B~) is stored.
    int out_rot;
};

template <typename T>
__device__ __forceinline__ long long line_offset(int pos, long long sline, int split, long long sblock) {
    return split ? (long long)(pos % split) * sline + (long long)(pos / split) * sblock : (long long)pos * sline;
}

template <typename T, int N, int DIR>
__global__ void __launch_bounds__(tx_for(N) * (N / 8)) k_lines(const LinesParams<T> p) {
    using C = cx_t<T>;
    constexpr int TX = tx_for(N);
    constexpr int TPL = N / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tx = threadIdx.x, j = threadIdx.y;
    const int kx = blockIdx.x * TX + tx;
    const bool valid = kx < p.ninner;
    LayoutRows<C, TX> lay(reinterpret_cast<C*>(smem_raw), tx);

    C v[8];
    const C* src = p.in + (long long)blockIdx.z * p.in_scomp + (long long)blockIdx.y * p.in_souter + kx;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int pos = j + q * TPL;
        if (p.in_fold) {
            const bool hi = pos > N / 2;
            const C t = valid ? __ldg(src + (long long)(hi ? N - pos : pos) * p.in_sline) : cmake<C>(0, 0);
            v[q] = hi ? cscale(t, p.fold_sign) : t;
        } else {
            v[q] = (valid && pos < p.n_in) ? ld_stream(src + line_offset<T>(pos, p.in_sline, p.in_split, p.in_sblock))
                                           : cmake<C>(0, 0);
        }
    }
    fft_line<T, N, DIR>(v, lay, j, p.tw);
    const int outer = p.use_peers == 2 ? (int)blockIdx.y % p.out_osplit : (int)blockIdx.y;
    const long long dst_off = (long long)blockIdx.z * p.out_scomp + (long long)outer * p.out_souter + kx;
    C* const outer_peer = p.use_peers == 2 ? p.out_peer[blockIdx.y / p.out_osplit] : nullptr;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int pos = j + q * TPL;
        if (valid && pos < p.n_out) {
            C o = cscale(v[q], p.scale);
            if (p.out_rot) o = cmake<C>(o.y, -o.x);
            if (p.use_peers == 2)
                outer_peer[dst_off + (long long)pos * p.out_sline] = o;
            else if (p.use_peers)
                p.out_peer[pos / p.out_split][dst_off + (long long)(pos % p.out_split) * p.out_sline] = o;
            else
                p.out[dst_off + line_offset<T>(pos, p.out_sline, p.out_split, p.out_sblock)] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// fused z pass
template <typename T>
struct ZParams {
    const cx_t<T>* in;      // B  [kx + PX*(ky + Ly*z)]
    cx_t<T>* out;           // C_c at out + c*out_scomp, same layout
    const cx_t<T>* tw;
    long long out_scomp;
    int nz;                 // valid z planes in and out
    int ncomp;              // 3: Ex,Ey,Ez; 4: + scalar potential (component 3, even Green function => real spectrum)
    int ninner, PX, Ly;     // Ly: ky lines held by this rank (= plane pitch / PX)
    int Lyg, ky0;           // global padded y length and first global ky of this rank (Lyg = Ly, ky0 = 0 on one GPU)
    // kx-slab solve (multi-GPU): this rank holds kx in [kx0, kx0 + ninner) with pitch PX; the Green spectrum keeps the
    // global extents ninner_g / pitch PXg.  One GPU and the ky-slab solve: kx0 = 0, ninner_g = ninner, PXg = PX.
    int kx0, ninner_g, PXg;
    // peer-memory output (multi-GPU): z plane pos goes to rank pos / out_split through out_peer[rank],
    // at local plane pos % out_split -- the all-to-all back is fused into the pass
    int use_peers, out_split;
    cx_t<T>* out_peer[SCB_MAX_RANKS];
    // GREEN_FREE / GREEN_CATHODE: compressed real spectrum S_c[kx + PXg*(ky' + (Ly/2+1)*kz')],
    // Green_c = i * sign * S_c with ky' = min(ky, Ly-ky), kz' = min(kz, Lz-kz)
    const T* S;
    long long S_scomp;
    // GREEN_CATHODE: image spectrum H_c, complex, folded like S: [kx + PX*(ky' + (Ly/2+1)*kz')] with
    //   H(ky > Ly/2) = p_y H(Ly-ky), H(kz > Lz/2) = p_x p_y conj(H(Lz-kz)), p = -1 along the component's axis
    // GREEN_FULL: G_c, complex, unfolded [kx + PX*(ky + Ly*kz)]
    const cx_t<T>* H;
    long long H_scomp;
    // k_z_eo: copy of S with kz' fastest, St[((c*ninner + kx)*(Ly/2+1) + ky')*PZ + kz']; cathode: the image spectrum
    // H in the same arrangement (complex entries)
    const T* St;
    const cx_t<T>* Ht;
    int PZ;
};

// cp.async of one real (4 or 8 bytes) from global to shared memory: lets the Green-spectrum
// values of the next field component travel while the inverse FFT of the current one runs,
// without holding them in registers (the kernel already sits at the 128-register limit).
template <typename T> __device__ __forceinline__ void cp_async_real(T* smem_dst, const T* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int N, int KIND>
__global__ void __launch_bounds__(tz_for(N) * (N / 8), z_minblocks<T>()) k_z_fused(const ZParams<T> p) {
    using C = cx_t<T>;
    constexpr int TX = tz_for(N);
    constexpr int TPL = N / 8;
    constexpr bool USE_S = (KIND == GREEN_FREE || KIND == GREEN_CATHODE);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tx = threadIdx.x, j = threadIdx.y;
    const int kx = blockIdx.x * TX + tx;
    // Lines ky and Ly-ky read the same folded Green-spectrum rows: schedule them back to back so the
    // second one finds them in L2 (single-GPU layout only; a rank's ky slab holds no such pairs).
    int kyl = blockIdx.y;              // line within this rank's ky slab
    if (p.ky0 == 0 && p.Ly == p.Lyg) {
        const int b = blockIdx.y, m = b >> 1;
        kyl = b == 0 ? 0 : b == 1 ? p.Ly / 2 : (b & 1) ? p.Ly - m : m;
    }
    const int ky = p.ky0 + kyl;        // global ky: selects the Green-spectrum entries
    const bool valid = kx < p.ninner;
    LayoutRows<C, TX> lay(reinterpret_cast<C*>(smem_raw), tx);
    // second buffer (cathode only) keeps the forward spectrum so that bin (-kz) can be read
    LayoutRows<C, TX> laym(reinterpret_cast<C*>(smem_raw + LayoutRows<C, TX>::bytes(N)), tx);
    // staging area for the compressed spectrum: slot (q, j, tx) is private to this thread
    T* sstage = reinterpret_cast<T*>(smem_raw + (KIND == GREEN_CATHODE ? 2 : 1) * LayoutRows<C, TX>::bytes(N)) +
                (j * TX + tx);
    constexpr int SSTRIDE = TPL * TX;
    const long long plane = (long long)p.PX * p.Ly;
    const int Lyh = p.Lyg / 2;
    const int kyf = ky <= Lyh ? ky : p.Lyg - ky;  // folded ky

    auto prefetch_S = [&](int c) {
        if constexpr (USE_S) {
            if (valid) {
                const T* base = p.S + c * p.S_scomp + (kx + p.kx0) + (long long)p.PXg * kyf;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int kz = j + q * TPL;
                    const int kzf = kz <= N / 2 ? kz : N - kz;
                    cp_async_real<T>(sstage + q * SSTRIDE, base + (long long)p.PXg * (Lyh + 1) * kzf);
                }
            }
            cp_async_commit();
        }
    };

    prefetch_S(0);
    C wr[TwN<N>::value];
    fft_twiddles<T, N>(wr, j, p.tw);
    C spec[8];
    {
        const C* src = p.in + (long long)kyl * p.PX + kx;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = j + q * TPL;
            spec[q] = (valid && pos < p.nz) ? ld_stream(src + pos * plane) : cmake<C>(0, 0);
        }
    }
    fft_line<T, N, -1>(spec, lay, j, wr);

    if constexpr (KIND == GREEN_CATHODE) {
#pragma unroll
        for (int q = 0; q < 8; ++q) laym.st(j + q * TPL, spec[q]);
        __syncthreads();
    }

#pragma unroll 1
    for (int c = 0; c < p.ncomp; ++c) {
        C w[8];
        if constexpr (USE_S) cp_async_wait_all();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int kz = j + q * TPL;
            C acc = cmake<C>(0, 0);
            if (valid) {
                if constexpr (USE_S) {
                    T s = sstage[q * SSTRIDE];
                    if ((c == 1 && ky > Lyh) || (c == 2 && kz > N / 2)) s = -s;
                    // field: (a + ib) * (i s) = s * (-b + i a);  potential: (a + ib) * s
                    acc = c == 3 ? cmake<C>(spec[q].x * s, spec[q].y * s) : cmake<C>(-spec[q].y * s, spec[q].x * s);
                }
                if constexpr (KIND == GREEN_CATHODE) {
                    const int kzf = kz <= N / 2 ? kz : N - kz;
                    C h = __ldg(p.H + c * p.H_scomp + (kx + p.kx0) + (long long)p.PXg * (kyf + (long long)(Lyh + 1) * kzf));
                    const T py = (c == 1) ? (T)-1 : (T)1, pxy = (c == 0 || c == 1) ? (T)-1 : (T)1;
                    if (ky > Lyh) h = cscale(h, py);
                    if (kz > N / 2) h = cmake<C>(pxy * h.x, -pxy * h.y);
                    const C m = laym.ld((N - kz) & (N - 1));
                    acc = cadd(acc, cmul(m, h));
                }
                if constexpr (KIND == GREEN_FULL) {
                    const C g = ld_stream(p.H + c * p.H_scomp + (kx + p.kx0) + (long long)p.PXg * (ky + (long long)p.Lyg * kz));
                    acc = cmul(spec[q], g);
                }
            }
            w[q] = acc;
        }
        if (c + 1 < p.ncomp) prefetch_S(c + 1);  // own slots only: no barrier needed before they are overwritten
        fft_line<T, N, +1>(w, lay, j, wr);
        const long long dst_off = c * p.out_scomp + (long long)kyl * p.PX + kx;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = j + q * TPL;
            if (valid && pos < p.nz) {
                if (p.use_peers) p.out_peer[pos / p.out_split][dst_off + (long long)(pos % p.out_split) * plane] = w[q];
                else p.out[dst_off + pos * plane] = w[q];
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// fused z pass, TMA variant (free space, single GPU, nz <= 256)
//
// ncu on k_z_fused (config 5, Float64): the LSU data pipe is the busiest unit (71 % of its wavefront rate): 72 % of
// the wavefronts are the shared-memory exchanges of the four transforms, the rest are the strided global loads and
// stores (64-byte segments) and the 8-byte cp.async prefetches of the Green spectrum, each of which costs a
// wavefront per 32-byte sector.  Here all global traffic goes through the TMA unit instead: one bulk tensor load
// brings the TX x nz input tile, two bring the folded Green-spectrum rows of a component (double-buffered one
// component ahead), and each component's TX x nz output tile leaves with one bulk tensor store from a staging
// buffer -- the LSU only carries the exchanges.  Out-of-range rows/columns are zero-filled on load and clipped on
// store by the tensor maps, so the padding costs no instructions.
//   mapB: rank 3 {2*PX, Ly, nz} of T, box {2*TX, 1, nz};  mapC: rank 4 {2*PX, Ly, nz, ncomp}, box {2*TX, 1, nz, 1};
//   mapS: rank 4 {PX, Ly/2+1, Lz/2+1, ncomp}, box {TX, 1, SR, 1}, SR = z_tma_srows (two boxes per component)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3, const void* src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(K) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ constexpr size_t round128(size_t b) { return (b + 127) / 128 * 128; }
// Box rows of the Green-spectrum tensor map for padded length N: two boxes cover kz' = 0..N/2, and a box must be a
// multiple of 128 bytes so that the second one lands on a 128-byte aligned shared-memory address (TMA requirement).
template <typename T> __host__ __device__ constexpr int z_tma_srows(int N) {
    const int row = tz_for(N) * (int)sizeof(T);
    const int q = row >= 128 ? 1 : 128 / row;
    return ((N / 2 + 2) / 2 + q - 1) / q * q;
}

template <typename T, int N>
struct ZTmaLayout {
    using C = cx_t<T>;
    static constexpr int TX = tz_for(N);
    static constexpr int SR = z_tma_srows<T>(N);                                // rows per Green-spectrum box
    static constexpr size_t EX = round128(LayoutRows<C, TX>::bytes(N));         // exchange buffer
    static constexpr size_t TILE = round128((size_t)(N / 2) * TX * sizeof(C));  // input tile / output staging
    static constexpr size_t SB = round128((size_t)2 * SR * TX * sizeof(T));     // one component's spectrum rows
    static constexpr size_t BARS = EX + 2 * TILE + 2 * SB;
    static constexpr size_t BYTES = BARS + 64;
};

template <typename T, int N>
__global__ void __launch_bounds__(tz_for(N) * (N / 8), z_minblocks<T>())
k_z_tma(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
        const __grid_constant__ CUtensorMap mapS, const ZParams<T> p) {
    using C = cx_t<T>;
    using LY = ZTmaLayout<T, N>;
    constexpr int TX = LY::TX;
    constexpr int TPL = N / 8;
    constexpr int SR = LY::SR;
    extern __shared__ __align__(128) unsigned char smem_tma[];
    const int tx = threadIdx.x, j = threadIdx.y;
    const bool leader = (tx == 0 && j == 0);
    const int kx0 = blockIdx.x * TX;
    int kyl = blockIdx.y;   // ky / Ly-ky back to back: they read the same folded spectrum rows (L2 reuse)
    {
        const int b = blockIdx.y, m = b >> 1;
        kyl = b == 0 ? 0 : b == 1 ? p.Ly / 2 : (b & 1) ? p.Ly - m : m;
    }
    const int ky = kyl;
    const int Lyh = p.Ly / 2;
    const int kyf = ky <= Lyh ? ky : p.Ly - ky;
    LayoutRows<C, TX> lay(reinterpret_cast<C*>(smem_tma), tx);
    auto tile = [&](int b) { return reinterpret_cast<C*>(smem_tma + LY::EX + (size_t)b * LY::TILE); };
    auto sbuf = [&](int b) { return reinterpret_cast<T*>(smem_tma + LY::EX + 2 * LY::TILE + (size_t)b * LY::SB); };
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_tma + LY::BARS);   // [0] input, [1],[2] spectrum
    const unsigned tile_bytes = (unsigned)p.nz * TX * sizeof(C);
    const unsigned s_bytes = 2u * SR * TX * sizeof(T);

    auto load_S = [&](int c) {   // leader only
        unsigned long long* bar = bars + 1 + (c & 1);
        mbar_expect_tx(bar, s_bytes);
        tma_load_4d(sbuf(c & 1), &mapS, kx0 + p.kx0, kyf, 0, c, bar);
        tma_load_4d(sbuf(c & 1) + SR * TX, &mapS, kx0 + p.kx0, kyf, SR, c, bar);
    };

    if (leader) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (leader) {
        mbar_expect_tx(bars + 0, tile_bytes);
        tma_load_3d(tile(0), &mapB, 2 * kx0, kyl, 0, bars + 0);
        load_S(0);
        if (p.ncomp > 1) load_S(1);
    }

    C wr[TwN<N>::value];
    fft_twiddles<T, N>(wr, j, p.tw);   // in flight while the input tile travels
    C spec[8];
    mbar_wait(bars + 0, 0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int pos = j + q * TPL;
        spec[q] = (q < 4 && pos < p.nz) ? tile(0)[pos * TX + tx] : cmake<C>(0, 0);
    }
    fft_line<T, N, -1>(spec, lay, j, wr);

#pragma unroll 1
    for (int c = 0; c < p.ncomp; ++c) {
        C w[8];
        mbar_wait(bars + 1 + (c & 1), (c >> 1) & 1);
        const T* sb = sbuf(c & 1) + tx;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int kz = j + q * TPL;
            const int kzf = kz <= N / 2 ? kz : N - kz;
            T s = sb[kzf * TX];
            if ((c == 1 && ky > Lyh) || (c == 2 && kz > N / 2)) s = -s;
            // field: (a + ib) * (i s) = s * (-b + i a);  potential: (a + ib) * s
            w[q] = c == 3 ? cmake<C>(spec[q].x * s, spec[q].y * s) : cmake<C>(-spec[q].y * s, spec[q].x * s);
        }
        fft_line<T, N, +1>(w, lay, j, wr);
        // staging buffer c&1 was handed to the store of component c-2 (c = 0 reuses the input tile, which every thread
        // finished reading before the first barrier of the forward transform)
        if (c >= 2 && leader) bulk_wait_read<1>();
        if (c >= 2 || N < 16) __syncthreads();
        C* ob = tile(c & 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int pos = j + q * TPL;
            if (pos < p.nz) ob[pos * TX + tx] = w[q];
        }
        fence_proxy_async();
        __syncthreads();
        if (leader) {
            tma_store_4d(&mapC, 2 * kx0, kyl, 0, c, ob);
            bulk_commit();
            if (c + 2 < p.ncomp) load_S(c + 2);   // everyone is past its reads of sbuf[c&1] (barrier above)
        }
    }
    if (leader) bulk_wait_read<0>();   // shared memory must outlive the stores that read it
}

// ------------------------------------------------------------------------------------------
// fused z pass, even/odd-bin variant (free space, single GPU, padded length 512)
//
// ncu on k_z_tma (config 5, Float64): the LSU data pipe (67 %) and the FP64 pipe (50 %) share the kernel, and ~90 % of
// the LSU wavefronts are the two shared-memory exchanges of each of the four 512-point transforms, separated by
// block-wide barriers.  Here the 512-point transform of the zero-padded line is never formed:
//     X[2k]   = FFT256(x)[k],   X[2k+1] = FFT256(x * w512^n)[k]                     (no arithmetic on the padding)
//     y[n<256] = IFFT256(W_even)[n] + w512^-n * IFFT256(W_odd)[n]                   (no arithmetic on discarded outputs)
// and each 256-point transform runs on 16 threads x 16 values as radix-16 x radix-16 with ONE exchange.  A line is
// owned by one warp: lanes 0-15 carry the even bins, lanes 16-31 the odd ones, so every exchange is warp-local
// (__syncwarp, no block barrier), the per-thread factor w512^(+-t) of the odd half folds into the inter-stage twiddle,
// and the two halves meet in one shuffle step.  Shared-memory traffic per line drops from 16 to 8 exchange halves.
// The Green spectrum is read from its kz'-fastest copy (one contiguous row per line and component, 1-D bulk copy);
// input and output tiles travel as bulk tensor copies with a hardware swizzle so that a warp reads / writes its column
// of the [z][kx] tile without bank conflicts.  tools/z_evenodd_model.py is the NumPy model of this schedule.
template <int DIR, typename C> __device__ __forceinline__ C cmul_c(C v, typename real_of<C>::type cr, typename real_of<C>::type ci_fwd) {
    using R = typename real_of<C>::type;
    const R ci = DIR < 0 ? ci_fwd : -ci_fwd;
    return cmake<C>(v.x * cr - v.y * ci, v.x * ci + v.y * cr);
}

// 16-point DFT, natural order in and out: 4 x 4 with the twiddles w16^(q0*k0) as literals
template <int DIR, typename C> __device__ __forceinline__ void dft16(C (&v)[16]) {
    using R = typename real_of<C>::type;
    const R c1 = (R)0.92387953251128675613, s1 = (R)0.38268343236508977173, hh = (R)0.70710678118654752440;
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0) dft4<DIR>(v[q0], v[q0 + 4], v[q0 + 8], v[q0 + 12]);
    v[5] = cmul_c<DIR>(v[5], c1, -s1);     // w16^1
    v[9] = cmul_c<DIR>(v[9], hh, -hh);     // w16^2
    v[13] = cmul_c<DIR>(v[13], s1, -c1);   // w16^3
    v[6] = cmul_c<DIR>(v[6], hh, -hh);     // w16^2
    v[10] = cmul_qturn<DIR>(v[10]);        // w16^4
    v[14] = cmul_c<DIR>(v[14], -hh, -hh);  // w16^6
    v[7] = cmul_c<DIR>(v[7], s1, -c1);     // w16^3
    v[11] = cmul_c<DIR>(v[11], -hh, -hh);  // w16^6
    v[15] = cmul_c<DIR>(v[15], -c1, s1);   // w16^9
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) dft4<DIR>(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);
    // X[k0 + 4*k1] sits in v[4*k0 + k1]: transpose the 4 x 4 index
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const C tmp = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = tmp;
        }
}

// v[k] *= p0 * b^k, k = 0..15, the powers formed in two interleaved chains (even / odd k)
template <typename C> __device__ __forceinline__ void apply_powers16(C (&v)[16], C p0, C b) {
    const C b2 = cmul(b, b);
    C pe = p0, po = cmul(p0, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[2 * k] = cmul(v[2 * k], pe);
        v[2 * k + 1] = cmul(v[2 * k + 1], po);
        if (k < 7) {
            pe = cmul(pe, b2);
            po = cmul(po, b2);
        }
    }
}

// 16 x 16 transpose among the 16 threads of a half-warp through its private buffer (pitch 17: conflict-free both ways)
template <typename C> __device__ __forceinline__ void exchange16(C (&v)[16], C* ex, int t) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) ex[k * 17 + t] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = ex[t * 17 + k];
}

// v[q] *= w32^(DIR * -q)... forward (DIR = -1): exp(-2 pi i q / 32); inverse: the conjugate
template <int DIR, typename C> __device__ __forceinline__ void mul_w32_powers(C (&v)[16]) {
    using R = typename real_of<C>::type;
    // cos / sin of q*pi/16, q = 1..7
    const R c[8] = {(R)1.0, (R)0.98078528040323044913, (R)0.92387953251128675613, (R)0.83146961230254523708,
                    (R)0.70710678118654752440, (R)0.55557023301960222474, (R)0.38268343236508977173, (R)0.19509032201612826785};
    const R s[8] = {(R)0.0, (R)0.19509032201612826785, (R)0.38268343236508977173, (R)0.55557023301960222474,
                    (R)0.70710678118654752440, (R)0.83146961230254523708, (R)0.92387953251128675613, (R)0.98078528040323044913};
#pragma unroll
    for (int q = 1; q < 8; ++q) v[q] = cmul_c<DIR>(v[q], c[q], -s[q]);
    v[8] = cmul_qturn<DIR>(v[8]);
#pragma unroll
    for (int q = 9; q < 16; ++q) v[q] = cmul_c<DIR>(v[q], -s[q - 8], -c[q - 8]);   // w32^(8+r) = -i * w32^r
}

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#ifndef SCB_ZEO_TX
#define SCB_ZEO_TX 4
#endif
#ifndef SCB_ZEO_MINBLOCKS
#define SCB_ZEO_MINBLOCKS 2
#endif
#ifndef SCB_ZEO_MINBLOCKS_F32
#define SCB_ZEO_MINBLOCKS_F32 4
#endif
template <typename T> __host__ __device__ constexpr int zeo_minblocks() { return sizeof(T) == 4 ? SCB_ZEO_MINBLOCKS_F32 : SCB_ZEO_MINBLOCKS; }

template <typename T>
struct ZEoLayout {
    using C = cx_t<T>;
    static constexpr int N = 512, M = 256, TX = SCB_ZEO_TX, PZ = 264;     // TX lines (= warps) per CTA
    static constexpr int ROWB = TX * (int)sizeof(C);                      // bytes per z row of the tile: 32 / 64 / 128
    static constexpr unsigned SWZ_MASK = ROWB / 16 - 1;                   // SWIZZLE_32B / _64B / _128B
    static constexpr size_t TILE = (size_t)M * ROWB;                      // one [z][kx] tile
    static constexpr size_t EX = (size_t)TX * 2 * 16 * 17 * sizeof(C);    // one exchange buffer per half-warp
    static constexpr size_t SB = (size_t)TX * 2 * PZ * sizeof(T);         // two spectrum rows per warp
    static constexpr size_t BARS = 3 * TILE + EX + SB;                    // two input tiles + output staging first
    static constexpr size_t BYTES = BARS + 128 + 1024;                    // barriers + slack for the 1 KB alignment
    // byte offset of element (z row n, column w) inside a swizzled tile (tile base 1 KB aligned)
    __device__ static __forceinline__ unsigned off(int n, int w) {
        const unsigned o = (unsigned)n * ROWB + (unsigned)w * (unsigned)sizeof(C);
        return o ^ (((o >> 7) & SWZ_MASK) << 4);
    }
};

// One [z][kx] tile per CTA (a persistent variant that prefetched the next input tiles measured slower: 1.22 vs 0.94 ms;
// the hardware's in-order CTA dispatch keeps neighbouring kx tiles, which share 128-byte lines, in flight together).
// Three tile buffers: [0] receives the input and, once every warp has taken its line out of it, serves as output
// staging for component 2; components 0 and 1 are staged in [1] and [2] -- so no component waits for the bulk store of
// the previous one to drain its buffer (a potential as fourth component reuses [1] after the first store has read it).
// CATH: the cathode image term is added in the spectral multiply, acc += R(kx, ky, -kz) * H(kx, ky, kz) (DESIGN.md
// section 3, identity (ii)).  The mirrored forward bin N - b of lane (h, t), register k2 sits in register 15 - k2 of
// lane (h, t') with t' = 15 - t (odd bins) or 16 - t (even bins, t > 0), and in the lane's own register (16 - k2) % 16
// for the even bins of t = 0: one shuffle per value, nothing parked in shared memory.  H comes from its kz'-fastest
// copy straight from global memory: for a fixed k2 the 32 lanes read 32 consecutive entries (512 bytes).
template <typename T, bool CATH>
__global__ void __launch_bounds__(32 * SCB_ZEO_TX, zeo_minblocks<T>())
k_z_eo(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC, const ZParams<T> p) {
    using C = cx_t<T>;
    using LY = ZEoLayout<T>;
    constexpr int N = LY::N, TX = LY::TX, PZ = LY::PZ;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int h = lane >> 4, t = lane & 15;
    const bool leader = threadIdx.x == 0;
    const int kx0 = blockIdx.x * TX, kx = kx0 + w;
    int kyl;   // ky / Ly-ky back to back: they read the same folded spectrum rows (L2 reuse)
    {
        const int b = blockIdx.y, m = b >> 1;
        kyl = b == 0 ? 0 : b == 1 ? p.Ly / 2 : (b & 1) ? p.Ly - m : m;
    }
    const int ky = kyl, Lyh = p.Ly / 2;
    const int kyf = ky <= Lyh ? ky : p.Ly - ky;
    const bool valid = kx < p.ninner;   // warp-uniform
    C* ex = reinterpret_cast<C*>(smem_raw + 3 * LY::TILE) + (w * 2 + h) * (16 * 17);
    T* sbuf = reinterpret_cast<T*>(smem_raw + 3 * LY::TILE + LY::EX) + (size_t)w * 2 * PZ;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + LY::BARS);   // [0] input, [1 + 2w + b] spectrum
    unsigned long long* sbar = bars + 1 + 2 * w;
    const unsigned tile_bytes = (unsigned)p.nz * LY::ROWB;
    const unsigned s_bytes = (unsigned)PZ * sizeof(T);

    auto load_S = [&](int c) {   // lane 0 of a valid warp
        const T* src = p.St + (((long long)c * p.ninner_g + kx + p.kx0) * (Lyh + 1) + kyf) * PZ;
        mbar_expect_tx(sbar + (c & 1), s_bytes);
        bulk_load_1d(sbuf + (c & 1) * PZ, src, s_bytes, sbar + (c & 1));
    };

    if (leader) {
#pragma unroll
        for (int i = 0; i < 1 + 2 * TX; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (leader) {
        mbar_expect_tx(bars + 0, tile_bytes);
        tma_load_3d(smem_raw, &mapB, 2 * kx0, kyl, 0, bars + 0);
    }
    if (valid && lane == 0) {
        load_S(0);
        if (p.ncomp > 1) load_S(1);
    }
    if constexpr (CATH) {
        // the image-spectrum rows of this line (one 4 KB row per component) are read straight from global memory inside
        // the component loop: ask the L2 for them now, while the input tile travels and the forward transform runs
        if (valid) {
            for (int c = 0; c < p.ncomp; ++c) {
                const char* row = reinterpret_cast<const char*>(p.Ht + (((long long)c * p.ninner_g + kx + p.kx0) * (Lyh + 1) + kyf) * PZ);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128 * lane));
                if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128 * 32));
            }
        }
    }
    // per-thread roots (table: exp(-2 pi i k / 512)): w256^t, w512^t (odd half), w512^(2t+h)
    const C bf = __ldg(p.tw + 2 * t);
    const C cf = h ? __ldg(p.tw + t) : cmake<C>(1, 0);
    const C bi = cconj(__ldg(p.tw + 2 * t + h));

    C spec[16];
    mbar_wait(bars + 0, 0);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + 16 * q;
        spec[q] = (valid && n < p.nz) ? *reinterpret_cast<const C*>(smem_raw + LY::off(n, w)) : cmake<C>(0, 0);
    }
    __syncthreads();   // every warp holds its line: buffer 0 may be overwritten by output staging from here on
#ifndef SCB_ZEO_NOCOMPUTE
    if (valid) {
        if (h) mul_w32_powers<-1>(spec);
        dft16<-1>(spec);
        apply_powers16(spec, cf, bf);
        exchange16(spec, ex, t);
        dft16<-1>(spec);   // spec[k2] = bin 2*(t + 16*k2) + h
    }
#endif

#pragma unroll 1
    for (int c = 0; c < p.ncomp; ++c) {
        C y[8];
        if (valid) {
            C wv[16];
            mbar_wait(sbar + (c & 1), (c >> 1) & 1);
            const T* sb = sbuf + (c & 1) * PZ;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const int b = 2 * (t + 16 * k2) + h;
                const int kzf = b <= N / 2 ? b : N - b;
                T s = sb[kzf];
                if ((c == 1 && ky > Lyh) || (c == 2 && b > N / 2)) s = -s;
                // field: (a + ib) * (i s) = s * (-b + i a);  potential: (a + ib) * s
                wv[k2] = c == 3 ? cmake<C>(spec[k2].x * s, spec[k2].y * s) : cmake<C>(-spec[k2].y * s, spec[k2].x * s);
                if constexpr (CATH) {
                    C hv = __ldg(p.Ht + (((long long)c * p.ninner_g + kx + p.kx0) * (Lyh + 1) + kyf) * PZ + kzf);
                    // folding rules of the image spectrum: H(Ly - ky) = p_y H(ky), H(Lz - kz) = p_x p_y conj(H(kz))
                    const T py = (c == 1) ? (T)-1 : (T)1, pxy = (c == 0 || c == 1) ? (T)-1 : (T)1;
                    if (ky > Lyh) hv = cscale(hv, py);
                    if (b > N / 2) hv = cmake<C>(pxy * hv.x, -pxy * hv.y);
                    const int srcl = h ? 31 - t : ((16 - t) & 15);
                    C m;
                    m.x = __shfl_sync(0xffffffffu, spec[15 - k2].x, srcl);
                    m.y = __shfl_sync(0xffffffffu, spec[15 - k2].y, srcl);
                    if (lane == 0) m = spec[(16 - k2) & 15];
                    wv[k2] = cadd(wv[k2], cmul(m, hv));
                }
            }
            __syncwarp();
            if (lane == 0 && c + 2 < p.ncomp) load_S(c + 2);   // the whole warp is past its reads of this row
#ifndef SCB_ZEO_NOCOMPUTE   // (timing experiment: memory traffic of the pass without its arithmetic)
            dft16<+1>(wv);
            apply_powers16(wv, cmake<C>(1, 0), bi);
            exchange16(wv, ex, t);
            dft16<+1>(wv);   // wv[n2] = half h of y[t + 16*n2]
            if (h) mul_w32_powers<+1>(wv);
#endif
            // even half + odd half: lanes 0-15 keep n2 = 0..7, lanes 16-31 keep n2 = 8..15
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const C snd = h ? wv[i] : wv[8 + i];
                const C keep = h ? wv[8 + i] : wv[i];
                C got;
                got.x = __shfl_xor_sync(0xffffffffu, snd.x, 16);
                got.y = __shfl_xor_sync(0xffffffffu, snd.y, 16);
                y[i] = h ? cadd(got, keep) : cadd(keep, got);   // even part + odd part on both halves
            }
        }
        unsigned char* stage = smem_raw + (size_t)((c + 1) % 3) * LY::TILE;
        if (c >= 3) {   // fourth component: its buffer was handed to the store of component c-3
            if (leader) bulk_wait_read<2>();
            __syncthreads();
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = t + 16 * (8 * h + i);
                if (n < p.nz) *reinterpret_cast<C*>(stage + LY::off(n, w)) = y[i];
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (leader) {
            tma_store_4d(&mapC, 2 * kx0, kyl, 0, c, stage);
            bulk_commit();
        }
    }
    if (leader) bulk_wait_read<0>();   // shared memory must outlive the stores that read it
}

// ------------------------------------------------------------------------------------------
// x passes: real <-> half-complex along the contiguous axis, two real lines per transform
// Generator for the x pass of the Green-spectrum build: the padded, wrap-around-placed IGF array is
// never materialised; element (X, Y, Z) is produced on the fly from the table D of differenced values
// (one per distinct displacement).  Same mapping as k_green_place in green.cu.
struct GreenGen {
    const double* D;      // null: the pass reads real lines from memory
    int n[3], L[3], sym[3], corr[3], dcnt[3];
    int icomp;
    int ly_lines;         // lines enumerate (Y, Z) with Y < ly_lines (L_y/2+1 when the y symmetry is exploited)
    double sign_all;
};

// displacement bookkeeping of one axis: returns false when the index lies in the zero band
__device__ __forceinline__ bool green_axis(const GreenGen& g, int a, int X, int& m0, double& sgn) {
    const int n = g.n[a], L = g.L[a];
    int d;
    if (g.corr[a]) {
        if (X > 2 * n - 2) return false;
        d = X - (n - 1);
    } else {
        if (X <= n - 1) d = X;
        else if (X >= L - (n - 1)) d = X - L;
        else return false;
    }
    if (g.sym[a]) {
        if (d < 0) { d = -d; if (a == g.icomp - 1) sgn = -sgn; }
        m0 = d;
    } else {
        m0 = d + n - 1;
    }
    return true;
}

template <typename T>
struct XParams {
    GreenGen gen;
    const void* in;
    void* out;
    const cx_t<T>* tw;
    long long nlines;       // real lines (ny*nz), complex lines on the other side match 1:1
    long long real_sline;   // elements between consecutive real lines
    int n_real;             // valid reals per line (nx): zero beyond on load, not stored on store
    int PX;                 // pitch of the complex lines
    long long real_scomp, cplx_scomp;  // blockIdx.y selects the field component
    T scale;
    // generator variant only: the generated lines are even (odd) about index 0, so their spectrum is purely real
    // (imaginary); real_out = 1 (2) stores just that part as a REAL line of pitch PX (columns ninner..PX-1 zeroed)
    int real_out;
    // kx-slab solve (multi-GPU): the complex line is cut into blocks of `split` bins, one per rank.
    //   k_x_r2c: bin k of line l is stored through out_peer[k / split] at (l + line0) * split + k % split (the pitch on the
    //            receiving side is `split`; line0 = this rank's first global line), + component * cplx_scomp
    //   k_x_c2r: bin k of line l is read from in + (k / split) * sblock + l * split + k % split
    int split;
    long long line0, sblock;
    void* out_peer[SCB_MAX_RANKS];
};

// SPLIT: kx-slab solve (multi-GPU), a compile-time switch so that the single-GPU instantiation carries none of it
template <typename T, int N, bool GEN, bool SPLIT = false>
__global__ void __launch_bounds__((N / 8) * lp_for(N)) k_x_r2c(const XParams<T> p) {
    using C = cx_t<T>;
    constexpr int TPL = N / 8;
    constexpr int ROW = LayoutLine<C>::row(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int j = threadIdx.x, lp = threadIdx.y;
    const long long pair = (long long)blockIdx.x * lp_for(N) + lp;
    const long long la = 2 * pair, lb = 2 * pair + 1;
    const bool va = la < p.nlines, vb = lb < p.nlines;
    LayoutLine<C> lay(reinterpret_cast<C*>(smem_raw) + (size_t)lp * ROW);

    C v[8];
    if constexpr (GEN) {
        // lines are (Y, Z) rows of the padded IGF array; resolve the y and z displacements once
        const GreenGen& g = p.gen;
        double sa = g.sign_all, sb = g.sign_all;
        int ya = 0, za = 0, yb = 0, zb = 0;
        bool oka = va, okb = vb;
        if (oka) oka = green_axis(g, 1, (int)(la % g.ly_lines), ya, sa) && green_axis(g, 2, (int)(la / g.ly_lines), za, sa);
        if (okb) okb = green_axis(g, 1, (int)(lb % g.ly_lines), yb, sb) && green_axis(g, 2, (int)(lb / g.ly_lines), zb, sb);
        const long long rowa = (long long)g.dcnt[0] * (ya + (long long)g.dcnt[1] * za);
        const long long rowb = (long long)g.dcnt[0] * (yb + (long long)g.dcnt[1] * zb);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = j + q * TPL;
            int mx = 0;
            double sx = 1.0;
            const bool okx = green_axis(g, 0, pos, mx, sx);
            const double a = (oka && okx) ? sa * sx * __ldg(g.D + rowa + mx) : 0.0;
            const double b = (okb && okx) ? sb * sx * __ldg(g.D + rowb + mx) : 0.0;
            v[q] = cmake<C>((T)a, (T)b);
        }
    } else {
        const T* ra = static_cast<const T*>(p.in) + (long long)blockIdx.y * p.real_scomp + la * p.real_sline;
        const T* rb = static_cast<const T*>(p.in) + (long long)blockIdx.y * p.real_scomp + lb * p.real_sline;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = j + q * TPL;
            const bool in_range = pos < p.n_real;
            const T a = (va && in_range) ? ld_stream(ra + pos) : (T)0;
            const T b = (vb && in_range) ? ld_stream(rb + pos) : (T)0;
            v[q] = cmake<C>(a, b);
        }
    }
    fft_line<T, N, -1>(v, lay, j, p.tw);

    // split the two interleaved Hermitian spectra: needs bins k and N-k
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) lay.st(j + q * TPL, v[q]);
    __syncthreads();
    C* oa = static_cast<C*>(p.out) + (long long)blockIdx.y * p.cplx_scomp + la * p.PX;
    C* ob = static_cast<C*>(p.out) + (long long)blockIdx.y * p.cplx_scomp + lb * p.PX;
    const T half = (T)0.5;
    if constexpr (GEN) {
        if (p.real_out) {
            T* ra = static_cast<T*>(p.out) + la * p.PX;
            T* rb = static_cast<T*>(p.out) + lb * p.PX;
            const bool re = p.real_out == 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = j + q * TPL;
                const C zk = v[q];
                const C zm = lay.ld((N - k) & (N - 1));
                if (va) ra[k] = re ? (zk.x + zm.x) * half : (zk.y - zm.y) * half;
                if (vb) rb[k] = re ? (zk.y + zm.y) * half : (zm.x - zk.x) * half;
            }
            // Nyquist bin (real for a real line; zero for an odd one), then zeros up to the pitch
            for (int k = N / 2 + j; k < p.PX; k += TPL) {
                const bool nyq = (k == N / 2) && re;   // thread 0 holds bin N/2 in v[4]
                if (va) ra[k] = nyq ? v[4].x : (T)0;
                if (vb) rb[k] = nyq ? v[4].y : (T)0;
            }
            return;
        }
    }
    if constexpr (SPLIT) {   // kx-slab solve: every bin goes to the rank that owns its kx block
        const long long coff = (long long)blockIdx.y * p.cplx_scomp;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = j + q * TPL;
            const C zk = v[q];
            const C zm = lay.ld((N - k) & (N - 1));
            const C A = cmake<C>((zk.x + zm.x) * half, (zk.y - zm.y) * half);
            const C B = cmake<C>((zk.y + zm.y) * half, (zm.x - zk.x) * half);
            C* dst = static_cast<C*>(p.out_peer[k / p.split]) + coff + (k % p.split);
            if (va) dst[(la + p.line0) * p.split] = A;
            if (vb) dst[(lb + p.line0) * p.split] = B;
        }
        if (j == 0) {
            const int k = N / 2;
            C* dst = static_cast<C*>(p.out_peer[k / p.split]) + coff + (k % p.split);
            if (va) dst[(la + p.line0) * p.split] = cmake<C>(v[4].x, 0);
            if (vb) dst[(lb + p.line0) * p.split] = cmake<C>(v[4].y, 0);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = j + q * TPL;
            const C zk = v[q];
            const C zm = lay.ld((N - k) & (N - 1));
            const C A = cmake<C>((zk.x + zm.x) * half, (zk.y - zm.y) * half);
            const C B = cmake<C>((zk.y + zm.y) * half, (zm.x - zk.x) * half);
            if (va) oa[k] = A;
            if (vb) ob[k] = B;
        }
        if (j == 0) {  // Nyquist bin N/2 lives in v[4] of thread 0
            if (va) oa[N / 2] = cmake<C>(v[4].x, 0);
            if (vb) ob[N / 2] = cmake<C>(v[4].y, 0);
        }
    }
}

template <typename T, int N, bool SPLIT = false>
__global__ void __launch_bounds__((N / 8) * lp_for(N)) k_x_c2r(const XParams<T> p) {
    using C = cx_t<T>;
    constexpr int TPL = N / 8;
    constexpr int ROW = LayoutLine<C>::row(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int j = threadIdx.x, lp = threadIdx.y;
    const long long pair = (long long)blockIdx.x * lp_for(N) + lp;
    const long long la = 2 * pair, lb = 2 * pair + 1;
    const bool va = la < p.nlines, vb = lb < p.nlines;
    LayoutLine<C> lay(reinterpret_cast<C*>(smem_raw) + (size_t)lp * ROW);
    // kx-slab solve: bin k sits in block k / split (one block per source rank), pitch `split` inside the block
    const long long lpitch = SPLIT ? p.split : p.PX;
    const C* ia = static_cast<const C*>(p.in) + (long long)blockIdx.y * p.cplx_scomp + la * lpitch;
    const C* ib = static_cast<const C*>(p.in) + (long long)blockIdx.y * p.cplx_scomp + lb * lpitch;
    auto bin = [&](int k) -> long long { return SPLIT ? (long long)(k / p.split) * p.sblock + (k % p.split) : (long long)k; };

    // Z[k] = A[k] + i B[k],  Z[N-k] = conj(A[k]) + i conj(B[k])
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = j + q * TPL;
        const C a = va ? ld_stream(ia + bin(k)) : cmake<C>(0, 0);
        const C b = vb ? ld_stream(ib + bin(k)) : cmake<C>(0, 0);
        if (k == 0) {
            lay.st(0, cmake<C>(a.x, b.x));
        } else {
            lay.st(k, cmake<C>(a.x - b.y, a.y + b.x));
            lay.st(N - k, cmake<C>(a.x + b.y, b.x - a.y));
        }
    }
    if (j == 0) {
        const C a = va ? ld_stream(ia + bin(N / 2)) : cmake<C>(0, 0);
        const C b = vb ? ld_stream(ib + bin(N / 2)) : cmake<C>(0, 0);
        lay.st(N / 2, cmake<C>(a.x, b.x));
    }
    __syncthreads();
    C v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = lay.ld(j + q * TPL);
    fft_line<T, N, +1>(v, lay, j, p.tw);

    T* oa = static_cast<T*>(p.out) + (long long)blockIdx.y * p.real_scomp + la * p.real_sline;
    T* ob = static_cast<T*>(p.out) + (long long)blockIdx.y * p.real_scomp + lb * p.real_sline;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int pos = j + q * TPL;
        if (pos < p.n_real) {
            if (va) oa[pos] = v[q].x * p.scale;
            if (vb) ob[pos] = v[q].y * p.scale;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers (defined per precision in fft_passes_f32.cu / fft_passes_f64.cu)
template <typename T> cudaError_t launch_lines(int N, int dir, const LinesParams<T>& p, int nouter, int ncomp, cudaStream_t s);
template <typename T> cudaError_t launch_z_fused(int N, int kind, const ZParams<T>& p, cudaStream_t s);
// TMA variant; returns cudaErrorNotSupported when this (N, nz) has no TMA instantiation
template <typename T> cudaError_t launch_z_tma(int N, const ZParams<T>& p, const CUtensorMap& mapB, const CUtensorMap& mapC,
                                                const CUtensorMap& mapS, cudaStream_t s);

template <typename T> cudaError_t launch_z_eo(const ZParams<T>& p, const CUtensorMap& mapB, const CUtensorMap& mapC, cudaStream_t s,
                                              bool cathode = false);
template <typename T> cudaError_t launch_x_r2c(int N, const XParams<T>& p, int ncomp, cudaStream_t s);
template <typename T> cudaError_t launch_x_c2r(int N, const XParams<T>& p, int ncomp, cudaStream_t s);
bool fft_len_supported(int N);

}  // namespace scb
