// ref_structure.cu -- SECONDARY GPU BASELINE (bench/test infrastructure, never part of the product).
//
// The reference's CUDA.jl path cannot run here (no Julia), so its *structure* is restated in plain
// CUDA + cuFFT, launch for launch as listed in SURVEY.md 2b (rows D1-I2):
//   deposit   : one thread per particle, eight global atomicAdd, block 256   (src/deposition.jl:106-158)
//   solve     : fill! + embed, in-place Z2Z forward, then per component {point-wise Green on the whole
//               (2n)^3 complex array, 8-point differencing into a temp, copy back, Z2Z forward, multiply,
//               Z2Z inverse, separate 1/M scaling pass, extract * FPEI}       (src/solvers/free_space.jl:56-101,
//               src/green_functions.jl:41-112); cathode = second solve on the flipped, negated rho + add
//   interpolate: one thread per particle, 24 scalar gathers                  (src/interpolation.jl:17-128)
// Float64 only (the reference's benchmarks are Float64).  It exists to answer "how much of the speed-up is the
// B200 and how much is the redesign": same GPU, reference structure vs. the library.
#include <cuda_runtime.h>
#include <cufft.h>
#include <math.h>
#include <stdint.h>

namespace {

constexpr double kFPEI = 299792458.0 * 299792458.0 * 1e-7;

struct Work {
    cufftHandle plan = 0;
    int n[3] = {0, 0, 0};
    cufftDoubleComplex *crho = nullptr, *cgrn = nullptr, *temp = nullptr;
    double *rho_img = nullptr, *e_img = nullptr;
};
Work g_w;

__global__ void k_deposit(long long np, const double* x, const double* y, const double* z, const double* q,
                          double* rho, int nx, int ny, double lx, double ly, double lz, double dx, double dy, double dz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double tx = (x[i] - lx) / dx, ty = (y[i] - ly) / dy, tz = (z[i] - lz) / dz;
    const long long ix = (long long)floor(tx), iy = (long long)floor(ty), iz = (long long)floor(tz);
    const double fx = tx - ix, fy = ty - iy, fz = tz - iz, c = q[i];
    double* r = rho + ix + (long long)nx * (iy + (long long)ny * iz);
    const long long sy = nx, sz = (long long)nx * ny;
    atomicAdd(r, c * (1 - fx) * (1 - fy) * (1 - fz));
    atomicAdd(r + 1, c * fx * (1 - fy) * (1 - fz));
    atomicAdd(r + sy, c * (1 - fx) * fy * (1 - fz));
    atomicAdd(r + sy + 1, c * fx * fy * (1 - fz));
    atomicAdd(r + sz, c * (1 - fx) * (1 - fy) * fz);
    atomicAdd(r + sz + 1, c * fx * (1 - fy) * fz);
    atomicAdd(r + sz + sy, c * (1 - fx) * fy * fz);
    atomicAdd(r + sz + sy + 1, c * fx * fy * fz);
}

__global__ void k_embed(cufftDoubleComplex* crho, const double* rho, int nx, int ny, int nz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nz) return;
    const int ix = i % nx, iy = (i / nx) % ny, iz = i / ((long long)nx * ny);
    crho[ix + 2LL * nx * (iy + 2LL * ny * iz)] = make_cuDoubleComplex(rho[i], 0.0);
}

__device__ double field_green(double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    return x * atan((y * z) / (r * x)) - z * log(r + y) + y * log((r - z) / (r + z)) / 2.0;
}

__global__ void k_green(cufftDoubleComplex* c, int sx, int sy, int sz, double dx, double dy, double dz0, double gamma,
                        int icomp, double ox, double oy, double oz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)sx * sy * sz) return;
    const int i = idx % sx, j = (idx / sx) % sy, k = idx / ((long long)sx * sy);
    const double dz = dz0 * gamma;
    const double factor = (icomp == 1 || icomp == 2) ? gamma / (dx * dy * dz) : 1.0 / (dx * dy * dz);
    const double u = i * dx + ((1 - sx) * dx / 2 + ox), v = j * dy + ((1 - sy) * dy / 2 + oy),
                 w = k * dz + ((1 - sz) * dz / 2 + oz * gamma);
    const double g = icomp == 1 ? field_green(u, v, w) : icomp == 2 ? field_green(v, w, u) : field_green(w, u, v);
    c[idx] = make_cuDoubleComplex(g * factor, 0.0);
}

__global__ void k_diff8(cufftDoubleComplex* out, const cufftDoubleComplex* c, int sx, int sy, int sz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)(sx - 1) * (sy - 1) * (sz - 1);
    if (idx >= tot) return;
    const int i = idx % (sx - 1), j = (idx / (sx - 1)) % (sy - 1), k = idx / ((long long)(sx - 1) * (sy - 1));
    auto at = [&](int a, int b, int d) { return c[a + (long long)sx * (b + (long long)sy * d)]; };
    cufftDoubleComplex r;
    r.x = at(i + 1, j + 1, k + 1).x - at(i, j + 1, k + 1).x - at(i + 1, j, k + 1).x - at(i + 1, j + 1, k).x - at(i, j, k).x +
          at(i, j, k + 1).x + at(i, j + 1, k).x + at(i + 1, j, k).x;
    r.y = at(i + 1, j + 1, k + 1).y - at(i, j + 1, k + 1).y - at(i + 1, j, k + 1).y - at(i + 1, j + 1, k).y - at(i, j, k).y +
          at(i, j, k + 1).y + at(i, j + 1, k).y + at(i + 1, j, k).y;
    out[i + (long long)sx * (j + (long long)sy * k)] = r;
}

__global__ void k_copy_back(cufftDoubleComplex* c, const cufftDoubleComplex* t, int sx, int sy, int sz) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)(sx - 1) * (sy - 1) * (sz - 1);
    if (idx >= tot) return;
    const int i = idx % (sx - 1), j = (idx / (sx - 1)) % (sy - 1), k = idx / ((long long)(sx - 1) * (sy - 1));
    const long long o = i + (long long)sx * (j + (long long)sy * k);
    c[o] = t[o];
}

__global__ void k_multiply(cufftDoubleComplex* t, const cufftDoubleComplex* a, const cufftDoubleComplex* b, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    t[i] = cuCmul(a[i], b[i]);
}

__global__ void k_scale(cufftDoubleComplex* t, double s, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    t[i].x *= s;
    t[i].y *= s;
}

__global__ void k_extract(double* e, const cufftDoubleComplex* t, int nx, int ny, int nz, double factr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nz) return;
    const int ix = i % nx, iy = (i / nx) % ny, iz = i / ((long long)nx * ny);
    e[i] = factr * t[(ix + nx - 1) + 2LL * nx * ((iy + ny - 1) + 2LL * ny * (iz + nz - 1))].x;
}

__global__ void k_flip_negate(double* out, const double* rho, int nx, int ny, int nz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nz) return;
    const int ix = i % nx, iy = (i / nx) % ny, iz = i / ((long long)nx * ny);
    out[i] = -rho[ix + (long long)nx * (iy + (long long)ny * (nz - 1 - iz))];
}

__global__ void k_add(double* a, const double* b, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] += b[i];
}

__global__ void k_interp(long long np, const double* x, const double* y, const double* z, const double* e, int nx, int ny,
                         int nz, double lx, double ly, double lz, double dx, double dy, double dz, double* ex, double* ey,
                         double* ez) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double tx = (x[i] - lx) / dx, ty = (y[i] - ly) / dy, tz = (z[i] - lz) / dz;
    const long long ix = (long long)floor(tx), iy = (long long)floor(ty), iz = (long long)floor(tz);
    const double fx = tx - ix, fy = ty - iy, fz = tz - iz;
    const double w[8] = {(1 - fx) * (1 - fy) * (1 - fz), fx * (1 - fy) * (1 - fz), (1 - fx) * fy * (1 - fz), fx * fy * (1 - fz),
                         (1 - fx) * (1 - fy) * fz,       fx * (1 - fy) * fz,       (1 - fx) * fy * fz,       fx * fy * fz};
    const long long sy = nx, sz = (long long)nx * ny, sc = sz * nz;
    const double* b = e + ix + sy * iy + sz * iz;
    double* out[3] = {ex, ey, ez};
    for (int c = 0; c < 3; ++c) {
        const double* bk = b + c * sc;
        out[c][i] = bk[0] * w[0] + bk[1] * w[1] + bk[sy] * w[2] + bk[sy + 1] * w[3] + bk[sz] * w[4] + bk[sz + 1] * w[5] +
                    bk[sz + sy] * w[6] + bk[sz + sy + 1] * w[7];
    }
}

inline unsigned nb(long long m) { return (unsigned)((m + 255) / 256); }

int ensure(const int n[3]) {
    if (g_w.n[0] == n[0] && g_w.n[1] == n[1] && g_w.n[2] == n[2]) return 0;
    if (g_w.plan) { cufftDestroy(g_w.plan); cudaFree(g_w.crho); cudaFree(g_w.cgrn); cudaFree(g_w.temp); cudaFree(g_w.rho_img); cudaFree(g_w.e_img); }
    const long long m = 8LL * n[0] * n[1] * n[2], ng = (long long)n[0] * n[1] * n[2];
    if (cudaMalloc(&g_w.crho, m * 16) || cudaMalloc(&g_w.cgrn, m * 16) || cudaMalloc(&g_w.temp, m * 16) ||
        cudaMalloc(&g_w.rho_img, ng * 8) || cudaMalloc(&g_w.e_img, 3 * ng * 8)) return -1;
    // Julia arrays are column-major: dims (2nx,2ny,2nz) with x fastest == cuFFT 3-D plan (2nz, 2ny, 2nx)
    if (cufftPlan3d(&g_w.plan, 2 * n[2], 2 * n[1], 2 * n[0], CUFFT_Z2Z) != CUFFT_SUCCESS) return -2;
    g_w.n[0] = n[0]; g_w.n[1] = n[1]; g_w.n[2] = n[2];
    return 0;
}

void solve_freespace(const double* rho, double* e, const int n[3], const double d[3], double gamma, const double off[3]) {
    const int sx = 2 * n[0], sy = 2 * n[1], sz = 2 * n[2];
    const long long m = (long long)sx * sy * sz, ng = (long long)n[0] * n[1] * n[2];
    cudaMemsetAsync(g_w.crho, 0, m * 16);
    k_embed<<<nb(ng), 256>>>(g_w.crho, rho, n[0], n[1], n[2]);
    cufftExecZ2Z(g_w.plan, g_w.crho, g_w.crho, CUFFT_FORWARD);
    for (int ic = 1; ic <= 3; ++ic) {
        k_green<<<nb(m), 256>>>(g_w.cgrn, sx, sy, sz, d[0], d[1], d[2], gamma, ic, off[0], off[1], off[2]);
        const long long md = (long long)(sx - 1) * (sy - 1) * (sz - 1);
        k_diff8<<<nb(md), 256>>>(g_w.temp, g_w.cgrn, sx, sy, sz);
        k_copy_back<<<nb(md), 256>>>(g_w.cgrn, g_w.temp, sx, sy, sz);
        cufftExecZ2Z(g_w.plan, g_w.cgrn, g_w.cgrn, CUFFT_FORWARD);
        k_multiply<<<nb(m), 256>>>(g_w.temp, g_w.crho, g_w.cgrn, m);
        cufftExecZ2Z(g_w.plan, g_w.temp, g_w.temp, CUFFT_INVERSE);
        k_scale<<<nb(m), 256>>>(g_w.temp, 1.0 / (double)m, m);
        k_extract<<<nb(ng), 256>>>(e + (ic - 1) * ng, g_w.temp, n[0], n[1], n[2], kFPEI);
    }
}

}  // namespace

extern "C" {

int refgpu_deposit(long long np, const double* x, const double* y, const double* z, const double* q, double* rho,
                   const int n[3], const double lo[3], const double d[3]) {
    cudaMemsetAsync(rho, 0, (size_t)n[0] * n[1] * n[2] * 8);
    k_deposit<<<nb(np), 256>>>(np, x, y, z, q, rho, n[0], n[1], lo[0], lo[1], lo[2], d[0], d[1], d[2]);
    return (int)cudaGetLastError();
}

int refgpu_solve(const double* rho, double* efield, const int n[3], const double lo[3], const double hi[3], const double d[3],
                 double gamma, int at_cathode) {
    if (ensure(n)) return -1;
    const double zero[3] = {0, 0, 0};
    solve_freespace(rho, efield, n, d, gamma, zero);
    if (at_cathode) {
        const long long ng = (long long)n[0] * n[1] * n[2];
        k_flip_negate<<<nb(ng), 256>>>(g_w.rho_img, rho, n[0], n[1], n[2]);
        const double off[3] = {0, 0, 2 * lo[2] + (hi[2] - lo[2])};
        solve_freespace(g_w.rho_img, g_w.e_img, n, d, gamma, off);
        k_add<<<nb(3 * ng), 256>>>(efield, g_w.e_img, 3 * ng);
    }
    return (int)cudaGetLastError();
}

int refgpu_interpolate(long long np, const double* x, const double* y, const double* z, const double* efield, const int n[3],
                       const double lo[3], const double d[3], double* ex, double* ey, double* ez) {
    k_interp<<<nb(np), 256>>>(np, x, y, z, efield, n[0], n[1], n[2], lo[0], lo[1], lo[2], d[0], d[1], d[2], ex, ey, ez);
    return (int)cudaGetLastError();
}

}  // extern "C"
