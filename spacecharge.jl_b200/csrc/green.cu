// green.cu -- integrated Green function (IGF) kernels, always evaluated in double.
//
// replaces: src/green_functions.jl:35-38 (field_green_function), :69-101 (get_green_kernel!),
// :103-112 (apply_8point_differencing!) and :41-67 (get_green_function!) of the reference.
//
// The reference fills the whole doubled (2n)^3 array point-wise and differences it, three times
// per solve.  Here the point-wise values are produced only on the corner ranges that are needed
// (an octant when the offset along an axis is zero: the IGF is odd along its own axis and even
// along the other two), differenced once per distinct displacement, and the zero-padded wrap-around
// array the FFT needs is generated inside the first FFT pass instead of being written to memory;
// the resulting spectrum is cached per geometry by the caller (api.cu).
//
// Compiled with -fmad=false so that u = (i-1)*dx + umin and the closed form round exactly like
// the reference's un-contracted Julia arithmetic.
#include "kernels.h"

namespace scb {

// src/green_functions.jl:35-38
__device__ __forceinline__ double field_green(double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    return x * atan((y * z) / (r * x)) - z * log(r + y) + y * log((r - z) / (r + z)) / 2.0;
}

// src/green_functions.jl:13-22 (defined by the reference but unreachable from its solve!; used here for
// the potential output, icomp = 0 -- SURVEY.md 8(f)-2)
__device__ __forceinline__ double potential_green(double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    if (r == 0.0) return 0.0;
    const double half = 0.5;
    return -half * (z * z) * atan(x * y / (z * r)) - half * (y * y) * atan(x * z / (y * r)) -
           half * (x * x) * atan(y * z / (x * r)) + y * z * log(x + r) + x * z * log(y + r) + x * y * log(z + r);
}

// Point-wise values P[mx + cx*(my + cy*mz)] for corner m <-> reference 1-based index i = i0 + m.
// src/green_functions.jl:69-101
__global__ void k_green_point(double* __restrict__ P, IgfGeom g, int icomp) {
    const long long total = (long long)g.cnt[0] * g.cnt[1] * g.cnt[2];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int mx = (int)(idx % g.cnt[0]);
    const int my = (int)((idx / g.cnt[0]) % g.cnt[1]);
    const int mz = (int)(idx / ((long long)g.cnt[0] * g.cnt[1]));
    const double dx = g.delta[0], dy = g.delta[1], dz = g.delta[2] * g.gamma;
    const double factor = (icomp == 1 || icomp == 2) ? g.gamma / (dx * dy * dz) : 1.0 / (dx * dy * dz);
    const double umin = (double)(1 - g.isize[0]) * dx / 2.0 + g.offset[0];
    const double vmin = (double)(1 - g.isize[1]) * dy / 2.0 + g.offset[1];
    const double wmin = (double)(1 - g.isize[2]) * dz / 2.0 + g.offset[2] * g.gamma;
    const double u = (double)(g.i0[0] + mx - 1) * dx + umin;
    const double v = (double)(g.i0[1] + my - 1) * dy + vmin;
    const double w = (double)(g.i0[2] + mz - 1) * dz + wmin;
    double gv;
    if (icomp == 1) gv = field_green(u, v, w) * factor;
    else if (icomp == 2) gv = field_green(v, w, u) * factor;
    else if (icomp == 3) gv = field_green(w, u, v) * factor;
    else if (icomp == 0) gv = potential_green(u, v, w) * factor;   // extension: scalar potential
    else gv = 0.0;
    P[idx] = gv;
}

// src/green_functions.jl:103-112, same left-to-right order
__device__ __forceinline__ double diff8(const double* __restrict__ P, int cx, int cy, int i, int j, int k) {
    auto at = [&](int a, int b, int c) { return __ldg(P + a + (long long)cx * (b + (long long)cy * c)); };
    return at(i + 1, j + 1, k + 1) - at(i, j + 1, k + 1) - at(i + 1, j, k + 1) - at(i + 1, j + 1, k) -
           at(i, j, k) + at(i, j, k + 1) + at(i, j + 1, k) + at(i + 1, j, k);
}

// Parity hook (scb_green): the reference's own array -- differenced block plus raw last planes.
template <typename TO>
__global__ void k_green_reference_layout(TO* __restrict__ out, const double* __restrict__ P, int sx, int sy, int sz) {
    const long long total = (long long)sx * sy * sz;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int i = (int)(idx % sx), j = (int)((idx / sx) % sy), k = (int)(idx / ((long long)sx * sy));
    double v;
    if (i < sx - 1 && j < sy - 1 && k < sz - 1) v = diff8(P, sx, sy, i, j, k);
    else v = P[idx];
    out[idx] = (TO)v;
}

// 8-point differencing on the stored corner ranges: D[a + dcx*(b + dcy*c)] = IGF for the cell whose
// lower corner is (a,b,c); one value per distinct displacement (an octant for symmetric axes).
__global__ void k_green_diff(double* __restrict__ D, const double* __restrict__ P, IgfGeom g) {
    const int dcx = g.cnt[0] - 1, dcy = g.cnt[1] - 1, dcz = g.cnt[2] - 1;
    const long long total = (long long)dcx * dcy * dcz;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int a = (int)(idx % dcx), b = (int)((idx / dcx) % dcy), c = (int)(idx / ((long long)dcx * dcy));
    D[idx] = diff8(P, g.cnt[0], g.cnt[1], a, b, c);
}

// The placement of the differenced values into the padded, wrap-around real array (displacement d at
// index d mod L, or d + n-1 along a correlation axis; symmetric axes read |d| with the parity sign)
// happens on the fly inside the x pass of the spectrum build: see GreenGen / green_axis in
// fft_passes.cuh.

// Spectrum -> cached forms.
// free space: Green_c = i*S_c, S_c real; the passes already pruned the spectrum to kx<=Lx/2,
// ky<=Ly/2, kz<=Lz/2 (layout [kx][ky][kz], pitch PX), so this only takes the imaginary part
template <typename T>
__global__ void k_green_compress_free(T* __restrict__ S, const double2* __restrict__ spec, int ninner, int PX, long long total,
                                      int take_real) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int kx = (int)(idx % PX);
    // field components: odd along one axis => purely imaginary spectrum; potential: even => purely real
    S[idx] = kx < ninner ? (T)(take_real ? spec[idx].x : spec[idx].y) : (T)0;
}

// S[kx + PX*(ky + Lyh1*kz)] -> St[(kx*Lyh1 + ky)*PZ + kz] (kz fastest): 32 x 32 tiles over (kx, kz) for each ky
template <typename T>
__global__ void __launch_bounds__(256) k_green_transpose(T* __restrict__ St, const T* __restrict__ S, int ninner, int PX, int Lyh1,
                                                          int Lzh1, int PZ) {
    __shared__ T tile[32][33];
    const int ky = blockIdx.z;
    const int kx0 = blockIdx.x * 32, kz0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int kz = kz0 + ty + 8 * r, kx = kx0 + tx;
        tile[ty + 8 * r][tx] = (kz < Lzh1 && kx < ninner) ? S[kx + (long long)PX * (ky + (long long)Lyh1 * kz)] : T{};
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int kx = kx0 + ty + 8 * r, kz = kz0 + tx;
        if (kx < ninner && kz < PZ) St[((long long)kx * Lyh1 + ky) * PZ + kz] = tile[tx][ty + 8 * r];
    }
}

// real-symmetric build, Float32 meshes: the passes run in double, the cached spectrum is Float32
__global__ void k_green_real_to_f32(float* __restrict__ S, const double* __restrict__ spec, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < total) S[idx] = (float)spec[idx];
}

template <typename T>
__global__ void k_green_convert_full(cx_t<T>* __restrict__ G, const double2* __restrict__ spec, int ninner, int PX, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int kx = (int)(idx % PX);
    double2 v = kx < ninner ? spec[idx] : make_double2(0.0, 0.0);
    G[idx] = cmake<cx_t<T>>((T)v.x, (T)v.y);
}

// ---- launchers -----------------------------------------------------------------------------
static inline unsigned blocks_for(long long total, int bs) { return (unsigned)((total + bs - 1) / bs); }

cudaError_t launch_green_point(double* P, const IgfGeom& g, int icomp, cudaStream_t s) {
    const long long total = (long long)g.cnt[0] * g.cnt[1] * g.cnt[2];
    k_green_point<<<blocks_for(total, 256), 256, 0, s>>>(P, g, icomp);
    return cudaGetLastError();
}

cudaError_t launch_green_reference_layout(void* out, int dt_f64, const double* P, int sx, int sy, int sz, cudaStream_t s) {
    const long long total = (long long)sx * sy * sz;
    if (dt_f64) k_green_reference_layout<double><<<blocks_for(total, 256), 256, 0, s>>>((double*)out, P, sx, sy, sz);
    else k_green_reference_layout<float><<<blocks_for(total, 256), 256, 0, s>>>((float*)out, P, sx, sy, sz);
    return cudaGetLastError();
}

cudaError_t launch_green_diff(double* D, const double* P, const IgfGeom& g, cudaStream_t s) {
    const long long nd = (long long)(g.cnt[0] - 1) * (g.cnt[1] - 1) * (g.cnt[2] - 1);
    k_green_diff<<<blocks_for(nd, 256), 256, 0, s>>>(D, P, g);
    return cudaGetLastError();
}

cudaError_t launch_green_transpose(void* St, const void* S, int dt_f64, int ninner, int PX, int Lyh1, int Lzh1, int PZ, cudaStream_t s,
                                   int complex_elems) {
    dim3 grid((ninner + 31) / 32, (PZ + 31) / 32, Lyh1);
    if (complex_elems) {   // image-charge spectrum: complex entries
        if (dt_f64) k_green_transpose<double2><<<grid, 256, 0, s>>>((double2*)St, (const double2*)S, ninner, PX, Lyh1, Lzh1, PZ);
        else k_green_transpose<float2><<<grid, 256, 0, s>>>((float2*)St, (const float2*)S, ninner, PX, Lyh1, Lzh1, PZ);
    } else if (dt_f64) k_green_transpose<double><<<grid, 256, 0, s>>>((double*)St, (const double*)S, ninner, PX, Lyh1, Lzh1, PZ);
    else k_green_transpose<float><<<grid, 256, 0, s>>>((float*)St, (const float*)S, ninner, PX, Lyh1, Lzh1, PZ);
    return cudaGetLastError();
}

cudaError_t launch_green_real_to_f32(void* S, const double* spec, long long total, cudaStream_t s) {
    k_green_real_to_f32<<<blocks_for(total, 256), 256, 0, s>>>((float*)S, spec, total);
    return cudaGetLastError();
}

cudaError_t launch_green_compress_free(void* S, int dt_f64, const double2* spec, int ninner, int PX, int Lyh1, int Lzh1,
                                       int take_real, cudaStream_t s) {
    const long long total = (long long)PX * Lyh1 * Lzh1;
    if (dt_f64) k_green_compress_free<double><<<blocks_for(total, 256), 256, 0, s>>>((double*)S, spec, ninner, PX, total, take_real);
    else k_green_compress_free<float><<<blocks_for(total, 256), 256, 0, s>>>((float*)S, spec, ninner, PX, total, take_real);
    return cudaGetLastError();
}

cudaError_t launch_green_convert_full(void* G, int dt_f64, const double2* spec, int ninner, int PX, long long total, cudaStream_t s) {
    if (dt_f64) k_green_convert_full<double><<<blocks_for(total, 256), 256, 0, s>>>((double2*)G, spec, ninner, PX, total);
    else k_green_convert_full<float><<<blocks_for(total, 256), 256, 0, s>>>((float2*)G, spec, ninner, PX, total);
    return cudaGetLastError();
}

}  // namespace scb
