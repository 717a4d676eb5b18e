// fft_passes_impl.cuh -- launcher bodies; included once per precision with SCB_T defined.
#include "fft_passes.cuh"

namespace scb {

#define SCB_FOR_EACH_N(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

template <> cudaError_t launch_lines<SCB_T>(int N, int dir, const LinesParams<SCB_T>& p, int nouter, int ncomp, cudaStream_t s) {
    using C = cx_t<SCB_T>;
    cudaError_t e = cudaErrorInvalidValue;
#define X(NN)                                                                                     \
    if (N == NN) {                                                                                \
        constexpr int TX = tx_for(NN);                                                            \
        const size_t sm = LayoutRows<C, TX>::bytes(NN);                                           \
        dim3 grid((p.ninner + TX - 1) / TX, nouter, ncomp), block(TX, NN / 8);                    \
        if (dir < 0) {                                                                            \
            e = set_smem(k_lines<SCB_T, NN, -1>, sm);                                             \
            if (e == cudaSuccess) k_lines<SCB_T, NN, -1><<<grid, block, sm, s>>>(p);              \
        } else {                                                                                  \
            e = set_smem(k_lines<SCB_T, NN, +1>, sm);                                             \
            if (e == cudaSuccess) k_lines<SCB_T, NN, +1><<<grid, block, sm, s>>>(p);              \
        }                                                                                         \
    }
    SCB_FOR_EACH_N(X)
#undef X
    return e != cudaSuccess ? e : cudaGetLastError();
}

// callers that do not use the kx-slab decomposition leave the spectrum-side extents at zero: same as the data's
static ZParams<SCB_T> with_spectrum_extents(ZParams<SCB_T> p) {
    if (p.PXg == 0) {
        p.PXg = p.PX;
        p.ninner_g = p.ninner;
    }
    return p;
}

template <> cudaError_t launch_z_fused<SCB_T>(int N, int kind, const ZParams<SCB_T>& p0, cudaStream_t s) {
    using C = cx_t<SCB_T>;
    const ZParams<SCB_T> p = with_spectrum_extents(p0);
    cudaError_t e = cudaErrorInvalidValue;
#define X(NN)                                                                                     \
    if (N == NN) {                                                                                \
        constexpr int TX = tz_for(NN);                                                            \
        const size_t sm1 = LayoutRows<C, TX>::bytes(NN);                                          \
        dim3 grid((p.ninner + TX - 1) / TX, p.Ly), block(TX, NN / 8);                             \
        const size_t sst = (size_t)TX * NN * sizeof(SCB_T);                                       \
        if (kind == GREEN_FREE) {                                                                 \
            e = set_smem(k_z_fused<SCB_T, NN, GREEN_FREE>, sm1 + sst);                            \
            if (e == cudaSuccess) k_z_fused<SCB_T, NN, GREEN_FREE><<<grid, block, sm1 + sst, s>>>(p); \
        } else if (kind == GREEN_CATHODE) {                                                       \
            e = set_smem(k_z_fused<SCB_T, NN, GREEN_CATHODE>, 2 * sm1 + sst);                     \
            if (e == cudaSuccess) k_z_fused<SCB_T, NN, GREEN_CATHODE><<<grid, block, 2 * sm1 + sst, s>>>(p); \
        } else {                                                                                  \
            e = set_smem(k_z_fused<SCB_T, NN, GREEN_FULL>, sm1);                                  \
            if (e == cudaSuccess) k_z_fused<SCB_T, NN, GREEN_FULL><<<grid, block, sm1, s>>>(p);   \
        }                                                                                         \
    }
    SCB_FOR_EACH_N(X)
#undef X
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <> cudaError_t launch_z_tma<SCB_T>(int N, const ZParams<SCB_T>& p0, const CUtensorMap& mapB, const CUtensorMap& mapC,
                                             const CUtensorMap& mapS, cudaStream_t s) {
    const ZParams<SCB_T> p = with_spectrum_extents(p0);
    cudaError_t e = cudaErrorNotSupported;
#define X(NN)                                                                                     \
    if (N == NN) {                                                                                \
        using LY = ZTmaLayout<SCB_T, NN>;                                                         \
        dim3 grid((p.ninner + LY::TX - 1) / LY::TX, p.Ly), block(LY::TX, NN / 8);                 \
        e = set_smem(k_z_tma<SCB_T, NN>, LY::BYTES);                                              \
        if (e == cudaSuccess) k_z_tma<SCB_T, NN><<<grid, block, LY::BYTES, s>>>(mapB, mapC, mapS, p); \
    }
    X(16) X(32) X(64) X(128) X(256) X(512)
#undef X
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <> cudaError_t launch_z_eo<SCB_T>(const ZParams<SCB_T>& p0, const CUtensorMap& mapB, const CUtensorMap& mapC, cudaStream_t s,
                                           bool cathode) {
    using LY = ZEoLayout<SCB_T>;
    const ZParams<SCB_T> p = with_spectrum_extents(p0);
    dim3 grid((p.ninner + LY::TX - 1) / LY::TX, p.Ly), block(32 * LY::TX);
    cudaError_t e;
    if (cathode) {
        e = set_smem(k_z_eo<SCB_T, true>, LY::BYTES);
        if (e == cudaSuccess) k_z_eo<SCB_T, true><<<grid, block, LY::BYTES, s>>>(mapB, mapC, p);
    } else {
        e = set_smem(k_z_eo<SCB_T, false>, LY::BYTES);
        if (e == cudaSuccess) k_z_eo<SCB_T, false><<<grid, block, LY::BYTES, s>>>(mapB, mapC, p);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <> cudaError_t launch_x_r2c<SCB_T>(int N, const XParams<SCB_T>& p, int ncomp, cudaStream_t s) {
    using C = cx_t<SCB_T>;
    cudaError_t e = cudaErrorInvalidValue;
#define X(NN)                                                                                     \
    if (N == NN) {                                                                                \
        constexpr int LP = lp_for(NN);                                                            \
        const size_t sm = (size_t)LP * LayoutLine<C>::row(NN) * sizeof(C);                        \
        const long long pairs = (p.nlines + 1) / 2;                                               \
        dim3 grid((unsigned)((pairs + LP - 1) / LP), ncomp), block(NN / 8, LP);                   \
        if (sizeof(SCB_T) == 8 && p.gen.D) {                                                      \
            e = set_smem(k_x_r2c<SCB_T, NN, sizeof(SCB_T) == 8>, sm);                             \
            if (e == cudaSuccess) k_x_r2c<SCB_T, NN, sizeof(SCB_T) == 8><<<grid, block, sm, s>>>(p); \
        } else if (p.split) {                                                                     \
            e = set_smem(k_x_r2c<SCB_T, NN, false, true>, sm);                                    \
            if (e == cudaSuccess) k_x_r2c<SCB_T, NN, false, true><<<grid, block, sm, s>>>(p);     \
        } else {                                                                                  \
            e = set_smem(k_x_r2c<SCB_T, NN, false>, sm);                                          \
            if (e == cudaSuccess) k_x_r2c<SCB_T, NN, false><<<grid, block, sm, s>>>(p);           \
        }                                                                                         \
    }
    SCB_FOR_EACH_N(X)
#undef X
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <> cudaError_t launch_x_c2r<SCB_T>(int N, const XParams<SCB_T>& p, int ncomp, cudaStream_t s) {
    using C = cx_t<SCB_T>;
    cudaError_t e = cudaErrorInvalidValue;
#define X(NN)                                                                                     \
    if (N == NN) {                                                                                \
        constexpr int LP = lp_for(NN);                                                            \
        const size_t sm = (size_t)LP * LayoutLine<C>::row(NN) * sizeof(C);                        \
        const long long pairs = (p.nlines + 1) / 2;                                               \
        dim3 grid((unsigned)((pairs + LP - 1) / LP), ncomp), block(NN / 8, LP);                   \
        if (p.split) {                                                                            \
            e = set_smem(k_x_c2r<SCB_T, NN, true>, sm);                                           \
            if (e == cudaSuccess) k_x_c2r<SCB_T, NN, true><<<grid, block, sm, s>>>(p);            \
        } else {                                                                                  \
            e = set_smem(k_x_c2r<SCB_T, NN>, sm);                                                 \
            if (e == cudaSuccess) k_x_c2r<SCB_T, NN><<<grid, block, sm, s>>>(p);                  \
        }                                                                                         \
    }
    SCB_FOR_EACH_N(X)
#undef X
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace scb
