/* ref_port.c -- C restatement of the reference's CPU hot path, for timing and for large oracle runs.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (see oracle/spacecharge_oracle.py header): used by tests/ to
 * cross-check the NumPy oracle and by bench.py's cpu_baseline / --impl reference legs.  The
 * product never links or calls it.  PARITY UNPINNED (no golden vectors in the reference, Julia
 * not installed): validated against the NumPy oracle, which carries the pins.
 *
 * Same algorithmic choices as SpaceCharge.jl v1.2.0 on CPU:
 *   - deposit is a serial loop over particles        (src/deposition.jl:167-197, 237-240)
 *   - Green fill, 8-point differencing, interpolation are threaded over the index range the
 *     way KernelAbstractions' CPU backend splits an ndrange   (src/green_functions.jl:52-61,
 *     src/interpolation.jl:114-125)
 *   - fill!/embed/multiply/extract are serial element-wise passes, as Julia broadcasts are
 *     (src/solvers/free_space.jl:68-69, 92, 98-99)
 * The FFTs are done by the caller (oracle/cpu_reference.py) with a threaded C2C FFT.
 * Arrays are column-major, complex arrays interleaved (re, im), all Float64.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* thread count of the threaded loops below.  Set explicitly by the caller: launchers such as torchrun export
 * OMP_NUM_THREADS=1 into the environment, which would silently serialise the baseline. */
void port_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int port_get_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* src/deposition.jl:28-86, serial driver :167-197 */
void port_deposit(int64_t np, const double* x, const double* y, const double* z, const double* q,
                  double* rho, const int64_t n[3], const double lo[3], const double d[3]) {
    const int64_t nx = n[0], ny = n[1];
    for (int64_t p = 0; p < np; ++p) {
        const double tx = (x[p] - lo[0]) / d[0], ty = (y[p] - lo[1]) / d[1], tz = (z[p] - lo[2]) / d[2];
        const int64_t ix = (int64_t)floor(tx), iy = (int64_t)floor(ty), iz = (int64_t)floor(tz);
        const double dx = tx - ix, dy = ty - iy, dz = tz - iz;
        const double wx0 = 1 - dx, wx1 = dx, wy0 = 1 - dy, wy1 = dy, wz0 = 1 - dz, wz1 = dz;
        const double c = q[p];
        double* r = rho + ix + nx * (iy + ny * iz);
        const int64_t sy = nx, sz = nx * ny;
        r[0] += c * wx0 * wy0 * wz0;
        r[1] += c * wx1 * wy0 * wz0;
        r[sy] += c * wx0 * wy1 * wz0;
        r[sy + 1] += c * wx1 * wy1 * wz0;
        r[sz] += c * wx0 * wy0 * wz1;
        r[sz + 1] += c * wx1 * wy0 * wz1;
        r[sz + sy] += c * wx0 * wy1 * wz1;
        r[sz + sy + 1] += c * wx1 * wy1 * wz1;
    }
}

/* src/green_functions.jl:35-38 */
static inline double field_green(double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    return x * atan((y * z) / (r * x)) - z * log(r + y) + y * log((r - z) / (r + z)) / 2;
}

/* get_green_kernel!, src/green_functions.jl:69-101: cgrn complex (2nx,2ny,2nz) */
void port_green_point(double* cgrn, const int64_t s[3], const double delta[3], double gamma, int icomp,
                      const double off[3]) {
    const double dx = delta[0], dy = delta[1], dz = delta[2] * gamma;
    const double factor = (icomp == 1 || icomp == 2) ? gamma / (dx * dy * dz) : 1.0 / (dx * dy * dz);
    const double umin = (double)(1 - s[0]) * dx / 2 + off[0];
    const double vmin = (double)(1 - s[1]) * dy / 2 + off[1];
    const double wmin = (double)(1 - s[2]) * dz / 2 + off[2] * gamma;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < s[2]; ++k) {
        const double w = (double)k * dz + wmin;
        for (int64_t j = 0; j < s[1]; ++j) {
            const double v = (double)j * dy + vmin;
            double* row = cgrn + 2 * (s[0] * (j + s[1] * k));
            for (int64_t i = 0; i < s[0]; ++i) {
                const double u = (double)i * dx + umin;
                double g;
                if (icomp == 1) g = field_green(u, v, w) * factor;
                else if (icomp == 2) g = field_green(v, w, u) * factor;
                else if (icomp == 3) g = field_green(w, u, v) * factor;
                else g = 0.0;
                row[2 * i] = g;
                row[2 * i + 1] = 0.0;
            }
        }
    }
}

/* apply_8point_differencing! into temp (src/green_functions.jl:103-112), complex arithmetic */
void port_diff8(double* out, const double* c, const int64_t s[3]) {
    const int64_t sx = s[0], sy = s[1], sz = s[2];
#define AT(i, j, k) (2 * ((i) + sx * ((j) + sy * (k))))
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < sz - 1; ++k)
        for (int64_t j = 0; j < sy - 1; ++j)
            for (int64_t i = 0; i < sx - 1; ++i)
                for (int t = 0; t < 2; ++t)
                    out[AT(i, j, k) + t] = c[AT(i + 1, j + 1, k + 1) + t] - c[AT(i, j + 1, k + 1) + t] -
                                           c[AT(i + 1, j, k + 1) + t] - c[AT(i + 1, j + 1, k) + t] -
                                           c[AT(i, j, k) + t] + c[AT(i, j, k + 1) + t] +
                                           c[AT(i, j + 1, k) + t] + c[AT(i + 1, j, k) + t];
}

/* cgrn[1:end-1,1:end-1,1:end-1] .= temp[...]  (src/green_functions.jl:64-66), serial broadcast */
void port_copy_back(double* c, const double* temp, const int64_t s[3]) {
    const int64_t sx = s[0], sy = s[1], sz = s[2];
    for (int64_t k = 0; k < sz - 1; ++k)
        for (int64_t j = 0; j < sy - 1; ++j)
            memcpy(c + AT(0, j, k), temp + AT(0, j, k), sizeof(double) * 2 * (sx - 1));
}

/* fill!(crho, 0); crho[1:nx,1:ny,1:nz] .= rho  (src/solvers/free_space.jl:68-69) */
void port_embed(double* crho, const double* rho, const int64_t n[3]) {
    const int64_t sx = 2 * n[0], sy = 2 * n[1], sz = 2 * n[2];
    memset(crho, 0, sizeof(double) * 2 * sx * sy * sz);
    for (int64_t k = 0; k < n[2]; ++k)
        for (int64_t j = 0; j < n[1]; ++j)
            for (int64_t i = 0; i < n[0]; ++i) crho[AT(i, j, k)] = rho[i + n[0] * (j + n[1] * k)];
}

/* @. temp = crho * cgrn  (src/solvers/free_space.jl:92) */
void port_multiply(double* temp, const double* a, const double* b, int64_t m) {
    for (int64_t i = 0; i < m; ++i) {
        const double ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        temp[2 * i] = ar * br - ai * bi;
        temp[2 * i + 1] = ar * bi + ai * br;
    }
}

/* efield[:,:,:,c] = factr * real(temp[nx:2nx-1, ...])  (src/solvers/free_space.jl:98-99) */
void port_extract(double* e, const double* temp, const int64_t n[3], double factr) {
    const int64_t sx = 2 * n[0], sy = 2 * n[1];
    for (int64_t k = 0; k < n[2]; ++k)
        for (int64_t j = 0; j < n[1]; ++j)
            for (int64_t i = 0; i < n[0]; ++i)
                e[i + n[0] * (j + n[1] * k)] = factr * temp[AT(i + n[0] - 1, j + n[1] - 1, k + n[2] - 1)];
}
#undef AT

/* interpolate_kernel!, src/interpolation.jl:17-86 */
void port_interpolate(int64_t np, const double* x, const double* y, const double* z, const double* e,
                      const int64_t n[3], const double lo[3], const double d[3], double* ex, double* ey, double* ez) {
    const int64_t sy = n[0], sz = n[0] * n[1], sc = n[0] * n[1] * n[2];
    double* out[3] = {ex, ey, ez};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < np; ++p) {
        const double tx = (x[p] - lo[0]) / d[0], ty = (y[p] - lo[1]) / d[1], tz = (z[p] - lo[2]) / d[2];
        const int64_t ix = (int64_t)floor(tx), iy = (int64_t)floor(ty), iz = (int64_t)floor(tz);
        const double dx = tx - ix, dy = ty - iy, dz = tz - iz;
        const double w000 = (1 - dx) * (1 - dy) * (1 - dz), w100 = dx * (1 - dy) * (1 - dz);
        const double w010 = (1 - dx) * dy * (1 - dz), w110 = dx * dy * (1 - dz);
        const double w001 = (1 - dx) * (1 - dy) * dz, w101 = dx * (1 - dy) * dz;
        const double w011 = (1 - dx) * dy * dz, w111 = dx * dy * dz;
        const double* b = e + ix + sy * iy + sz * iz;
        for (int c = 0; c < 3; ++c) {
            const double* bk = b + c * sc;
            out[c][p] = bk[0] * w000 + bk[1] * w100 + bk[sy] * w010 + bk[sy + 1] * w110 + bk[sz] * w001 +
                        bk[sz + 1] * w101 + bk[sz + sy] * w011 + bk[sz + sy + 1] * w111;
        }
    }
}
