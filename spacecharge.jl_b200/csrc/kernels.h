// kernels.h -- internal interface between the kernel translation units and api.cu.
#pragma once
#include "common.cuh"

namespace scb {

struct IgfGeom {
    int n[3];       // grid size
    int L[3];       // padded FFT length per axis (power of two >= 2n)
    int isize[3];   // the reference's doubled array size 2n (src/green_functions.jl:71)
    int i0[3];      // first 1-based reference index held in the point-wise table
    int cnt[3];     // corners held per axis
    int sym[3];     // 1: octant stored (offset == 0 along the axis), mirrored with parity
    int corr[3];    // 1: correlation placement (image charge), 0: wrap-around convolution
    double delta[3];
    double gamma;
    double offset[3];
};

struct Geom3 {
    double lo[3];
    double delta[3];
    double rinv[3];   // RN(1 / delta), formed on the host (div_exact in particle_common.cuh)
    float rinvf[3];   // RN_f32(1 / (float)delta) for the all-Float32 kernels
    int n[3];
    int l2_keep = 1;   // 1: grid-side accesses of the particle passes carry an L2 evict_last policy (SCB_L2_HINT=0 disables)
    // gather kernels only: handle the particles whose z cell lies in [zlo, zhi) but not in [exlo, exhi) -- the
    // slab-decomposed step interpolates while the field slabs are still arriving.  zfirst: read z alone, fetch x and y
    // only for the selected particles (passes that select few of them).
    int zlo = 0, zhi = 0x7fffffff, exlo = 0, exhi = 0, zfirst = 0;
};

// fused momentum kick of the gather kernels: off = plain interpolate_field (outputs overwritten)
struct Kick {
    int on = 0;
    double cxy = 0.0, cz = 0.0;   // p_{x,y} += cxy * E_{x,y},  p_z += cz * E_z
};

// element strides of the caller's particle arrays (scb_*_strided entry points, SURVEY.md 8(f)-3: AoS / strided
// layouts of Bmad-style callers); contiguous = all 1.  q = 0 means one charge shared by every particle.
struct PLayout {
    long long x = 1, y = 1, z = 1, q = 1;   // inputs
    long long ex = 1, ey = 1, ez = 1;       // outputs (interpolated field, or the momenta of the fused kick)
};

// green.cu
cudaError_t launch_green_point(double* P, const IgfGeom& g, int icomp, cudaStream_t s);
cudaError_t launch_green_reference_layout(void* out, int dt_f64, const double* P, int sx, int sy, int sz, cudaStream_t s);
// D: the differenced values, (cnt-1)^3 doubles
cudaError_t launch_green_diff(double* D, const double* P, const IgfGeom& g, cudaStream_t s);
// complex_elems: the entries are complex (image-charge spectrum) instead of real (free-space S)
cudaError_t launch_green_transpose(void* St, const void* S, int dt_f64, int ninner, int PX, int Lyh1, int Lzh1, int PZ, cudaStream_t s,
                                   int complex_elems = 0);
cudaError_t launch_green_real_to_f32(void* S, const double* spec, long long total, cudaStream_t s);
cudaError_t launch_green_compress_free(void* S, int dt_f64, const double2* spec, int ninner, int PX, int Lyh1, int Lzh1,
                                       int take_real, cudaStream_t s);
cudaError_t launch_green_convert_full(void* G, int dt_f64, const double2* spec, int ninner, int PX, long long total, cudaStream_t s);

// particles.cu  (pdt/mdt: 0 = f32, 1 = f64)
// mode 1: one thread per particle; otherwise two lanes per particle (x-neighbours coalesce in L2)
cudaError_t launch_deposit(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                           const void* q, void* rho, const Geom3& g, int mode, cudaStream_t s,
                           const PLayout* lay = nullptr);
// cell-tile deposit: `tiles` = 4 * Ng mesh elements of scratch; zeroes it, accumulates, folds into rho
// (rho is overwritten, or added to when accumulate != 0)
cudaError_t launch_deposit_tiles(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                 const void* q, void* tiles, void* rho, const Geom3& g, int accumulate, cudaStream_t s,
                                 const PLayout* lay = nullptr);
cudaError_t launch_interpolate(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                               const void* efield, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                               const Kick& kick = Kick(), const PLayout* lay = nullptr);
// node-major repack of efield (32 bytes per node) and the gather that reads it
size_t packed_bytes_per_node(int mdt);   // node-major record size
int interp_mode();   // SCB_INTERP_MODE: 0 = default gather kernels
cudaError_t launch_pack_efield(int mdt, const void* efield, void* packed, const Geom3& g, cudaStream_t s,
                               long long first_node = 0, long long count = -1);
cudaError_t launch_interpolate_packed(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                      const void* packed, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                                      const Kick& kick = Kick(), const PLayout* lay = nullptr);
cudaError_t launch_bfield(int mdt, const void* efield, void* bfield, long long ng, double beta_over_c, cudaStream_t s);
cudaError_t launch_cell_index(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                              const Geom3& g, long long* ix, long long* iy, long long* iz, cudaStream_t s);
// partial[0..2] = min, partial[3..5] = max as doubles; must be initialised by the launcher
cudaError_t launch_bounds(int pdt, long long np, const void* x, const void* y, const void* z,
                          double* out6, cudaStream_t s, const PLayout* lay = nullptr);

// sorted.cu -- cell-ordered bunches: keys, radix sort of (key, index), permutation, run-accumulating deposit and gather
// scratch layout: [keys0 | keys1 | values | histograms | partials]; launch_cell_keys fills keys0 and the first pass's
// histogram table, launch_sort_pairs does the rest
cudaError_t launch_cell_keys(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const Geom3& g,
                             void* scratch, int key_bits, cudaStream_t s, const PLayout* lay = nullptr);
size_t sort_scratch_bytes(long long n);
cudaError_t launch_sort_pairs(void* scratch, long long n, int key_bits, unsigned* perm_out, cudaStream_t s, int* launches);
cudaError_t launch_permute(int elem_bytes, long long n, const unsigned* perm, int nf, const void* const* src,
                           void* const* dst, cudaStream_t s);
// counters2[0] = sampled neighbour pairs in the same or the x-adjacent cell, counters2[1] = sampled pairs
cudaError_t launch_order_probe(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const Geom3& g,
                               unsigned long long* counters2, cudaStream_t s);
// rho must have been cleared (or hold the grid to accumulate into); any particle order is correct.  tile: keep a
// shared-memory tile of rho under the walk (k_deposit_window; measured slower than the plain walk, kept for comparison)
cudaError_t launch_deposit_runs(int pdt, int mdt, long long np, const void* x, const void* y, const void* z, const void* q,
                                void* rho, const Geom3& g, cudaStream_t s, bool tile = false);
// gathers straight from efield (no node-major copy)
cudaError_t launch_interpolate_runs(int pdt, int mdt, long long np, const void* x, const void* y, const void* z,
                                    const void* efield, const Geom3& g, void* ex, void* ey, void* ez, cudaStream_t s,
                                    const Kick& kick = Kick());

// measurement only: mode 0 = random 256-bit sector reads, mode 1 = random four-lane fp64 sector reductions over
// `bytes` of `buf`; *ops receives the number of 32-byte sector operations the launch performs
cudaError_t launch_probe(int mode, void* buf, size_t bytes, int iters, unsigned grid, cudaStream_t s, double* ops);

}  // namespace scb
