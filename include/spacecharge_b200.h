/* spacecharge_b200.h -- C ABI of the B200-native space-charge hot path.
 *
 * Drop-in boundary for bmad-sim/SpaceCharge.jl v1.2.0 (paths below are relative to the
 * reference root).  The reference has no FFI of its own: its boundary is the exported Julia
 * surface `Mesh3D, deposit!, clear_mesh!, interpolate_field, solve!` (src/SpaceCharge.jl:17).
 * A Julia shim (spacecharge.jl_b200/julia/SpaceChargeB200.jl) keeps that surface and `ccall`s
 * the entry points declared here; the Python mirror (spacecharge.jl_b200/__init__.py) binds the
 * same symbols through ctypes and is what the tests and bench.py drive.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an int: 0 = SCB_OK, negative = scb_status error code; nothing
 *     throws or exits.  scb_last_error(h) returns a human-readable message for the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller (never retained past the call)
 *     unless the name says `_host`.  Work is enqueued asynchronously on the handle's stream;
 *     call scb_sync() (or synchronise the stream yourself) before reading results.
 *   - array layout is the reference's (Julia column-major): rho[ix + nx*(iy + ny*iz)],
 *     efield[ix + nx*(iy + ny*(iz + nz*c))], c = 0,1,2; particles are separate x,y,z,q arrays.
 *   - `pdt` is the element type of the particle arrays, `mdt` of the mesh arrays
 *     (scb_dtype).  Arithmetic happens in promote(pdt, mdt) exactly as Julia promotes
 *     (src/deposition.jl:39-41, src/interpolation.jl:31-33).
 *   - geometry (min_bounds, max_bounds, delta, gamma, offset) is passed as double holding the
 *     exact mdt-valued numbers stored in the Mesh3D struct (src/mesh.jl:19-34).
 *   - a handle is bound to one device and one stream and is not thread-safe.
 */
#ifndef SPACECHARGE_B200_H
#define SPACECHARGE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SCB_API __attribute__((visibility("default")))
#else
#define SCB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scb_handle scb_handle;

typedef enum scb_dtype { SCB_F32 = 0, SCB_F64 = 1 } scb_dtype;

typedef enum scb_status {
    SCB_OK = 0,
    SCB_ERR_INVALID_ARG = -1,   /* null pointer, bad dtype, grid dim < 2, np < 0 ...            */
    SCB_ERR_UNSUPPORTED = -2,   /* grid dimension beyond the supported padded FFT length        */
    SCB_ERR_CUDA = -3,          /* a CUDA runtime call or kernel launch failed                  */
    SCB_ERR_NO_DEVICE = -4,     /* no usable sm_100 device: there is NO CPU fallback            */
    SCB_ERR_ALLOC = -5,         /* device workspace allocation failed                           */
    SCB_ERR_COMM = -6           /* multi-GPU communicator failure                               */
} scb_status;

/* Library-level knobs (all optional; pass NULL for defaults). */
typedef struct scb_options {
    int32_t green_cache;       /* 1 (default): keep the IGF spectrum per geometry; 0: rebuild every
                                  solve like the reference does (src/solvers/free_space.jl:79-89) */
    int32_t deposit_mode;      /* 0 = auto, 1 = one thread per particle, 2 = lane pairs, 3 = cell tiles  */
    int32_t particle_order;    /* scb_particle_order: 0 = SCB_ORDER_RANDOM (default), 1 = SCB_ORDER_CELL; see
                                  scb_set_particle_order                                              */
    int32_t reserved[5];
} scb_options;

/* Per-stage device times of the most recent calls, in milliseconds (CUDA events on the
 * handle's stream; valid after scb_sync).  Enabled with scb_enable_timing(h, 1). */
typedef struct scb_timing {
    float deposit_ms;
    float solve_ms;
    float interpolate_ms;
    float green_ms;            /* time spent (re)building the IGF spectrum inside solve_ms        */
    float pass_ms[8];          /* F1, F2, Z, B2, B3 of the last solve; slab-decomposed solve: [0] includes the
                                  reduce-scatter of rho, [4] the all-gather of E, reported on their own as
                                  [5] and [6] ([7] reserved)                                          */
} scb_timing;

/* ---- lifecycle ------------------------------------------------------------------------- */
SCB_API int scb_version(void);
SCB_API int scb_create(int device, void* cuda_stream, const scb_options* opt, scb_handle** out);
SCB_API int scb_destroy(scb_handle* h);
SCB_API int scb_set_stream(scb_handle* h, void* cuda_stream);
SCB_API int scb_sync(scb_handle* h);
SCB_API const char* scb_last_error(const scb_handle* h);
SCB_API int scb_enable_timing(scb_handle* h, int on);
SCB_API int scb_get_timing(scb_handle* h, scb_timing* out);
/* number of kernels this library has launched through the handle since creation */
SCB_API int64_t scb_launch_count(const scb_handle* h);

/* ---- a5: clear_mesh!  (src/deposition.jl:10-12) ------------------------------------------ */
SCB_API int scb_clear(scb_handle* h, void* rho, const int64_t n[3], int mdt);

/* ---- a6/a7: deposit!  (src/deposition.jl:28-86, 218-247) --------------------------------- */
SCB_API int scb_deposit(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                const void* q, int pdt, void* rho, int mdt, const int64_t n[3],
                const double min_bounds[3], const double delta[3], int clear);

/* ---- a12/a13: solve!, solve_freespace!  (src/solvers/free_space.jl:14-47, 56-101) -------- */
SCB_API int scb_solve(scb_handle* h, const void* rho, void* efield, int mdt, const int64_t n[3],
              const double min_bounds[3], const double max_bounds[3], const double delta[3],
              double gamma, int at_cathode);
SCB_API int scb_solve_freespace(scb_handle* h, const void* rho, void* efield, int mdt,
                        const int64_t n[3], const double delta[3], double gamma,
                        const double offset[3]);

/* ---- extension: scalar potential and magnetic field ---------------------- */
/* The reference stores only rho and efield (src/mesh.jl:19-34); its potential_green_function
 * (src/green_functions.jl:13-22) is defined but unreachable because get_green_kernel! returns zero for
 * icomp outside 1..3 (src/green_functions.jl:90-98).  scb_solve_potential runs the same convolution
 * with that function as a fourth component (icomp = 0 with factor 1/(dx*dy*dz), the convention of the
 * upstream OpenSpaceCharge code the reference README cites) and writes phi (nx,ny,nz) next to efield:
 *     phi[p] = FPEI * sum_n rho[n] * IGF_phi((p - n) * delta [+ image term when at_cathode])
 * i.e. the potential in the bunch rest frame (dz = delta_z * gamma), so that E_{x,y} = -gamma * d(phi)/d{x,y}
 * and E_z = -(1/gamma) * d(phi)/dz on the lab-frame mesh.  efield is written exactly as by scb_solve. */
SCB_API int scb_solve_potential(scb_handle* h, const void* rho, void* efield, void* phi, int mdt,
                                const int64_t n[3], const double min_bounds[3], const double max_bounds[3],
                                const double delta[3], double gamma, int at_cathode);
/* Magnetic field of a bunch moving along +z with the mesh's gamma: B = (beta/c) z_hat x E, i.e.
 * Bx = -(beta/c) Ey, By = (beta/c) Ex, Bz = 0; bfield has the layout of efield. */
SCB_API int scb_bfield(scb_handle* h, const void* efield, void* bfield, int mdt, const int64_t n[3],
                       double gamma);

/* ---- a14: interpolate_field  (src/interpolation.jl:17-128) -------------------------------- */
SCB_API int scb_interpolate(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                    int pdt, const void* efield, int mdt, const int64_t n[3],
                    const double min_bounds[3], const double delta[3],
                    void* ex, void* ey, void* ez);

/* ---- extension: interpolation fused with the momentum kick ---------------- */
/* Same gather as scb_interpolate, but instead of returning E it updates the caller's momentum
 * arrays in place: px += coef_xy*Ex, py += coef_xy*Ey, pz += coef_z*Ez (product and sum formed
 * separately in promote(pdt, mdt), rounded to pdt once) -- E never round-trips through memory.
 * For a bunch moving along z the Lorentz force q(E + v x B) gives coef_xy = q*dt/gamma^2, coef_z = q*dt. */
SCB_API int scb_interpolate_kick(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                                 int pdt, const void* efield, int mdt, const int64_t n[3],
                                 const double min_bounds[3], const double delta[3],
                                 void* px, void* py, void* pz, double coef_xy, double coef_z);

/* ---- a8-a11: get_green_function!  (src/green_functions.jl:41-112), parity hook ----------- */
/* Writes the integrated Green function in the reference's layout: a REAL array of shape
 * n2 = (2nx, 2ny, 2nz) (the real part of the reference's complex cgrn; its imaginary part is
 * zero), entry (i,j,k) for displacement (i+1-nx, j+1-ny, k+1-nz); the last plane of every
 * dimension holds the raw point-wise values exactly like the reference leaves them.
 * Always evaluated in double; `dt` selects the output element type.  icomp = 1,2,3 as in the
 * reference; icomp = 0 gives the potential Green function used by scb_solve_potential. */
SCB_API int scb_green(scb_handle* h, void* cgrn_real_out, const int64_t n2[3], const double delta[3],
              double gamma, int icomp, const double offset[3], int dt);

/* ---- a2: extrema for the auto-bounds constructor  (src/mesh.jl:120-122) ------------------ */
/* Synchronous: returns host doubles holding the exact pdt-valued extrema. */
SCB_API int scb_bounds(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
               int pdt, double out_min[3], double out_max[3]);

/* ---- parity hook: particle -> cell indices (A.2), int64 outputs, unclamped ---------------- */
SCB_API int scb_cell_index(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                   int pdt, int mdt, const double min_bounds[3], const double delta[3],
                   int64_t* ix, int64_t* iy, int64_t* iz);

/* ---- fused step on device-resident data: deposit! + solve! + interpolate_field ------------ */
/* (the timed body of benchmark/full_pipeline_benchmark.jl:26-30) */
SCB_API int scb_step(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
             const void* q, int pdt, void* rho, void* efield, int mdt, const int64_t n[3],
             const double min_bounds[3], const double max_bounds[3], const double delta[3],
             double gamma, int at_cathode, void* ex, void* ey, void* ez);

/* ---- extension: strided / array-of-structures particle layouts ------------ */
/* The reference takes four separate dense vectors (src/deposition.jl:218-223,
 * src/interpolation.jl:100-104).  Tracking codes such as Bmad keep one record per particle (the
 * phase-space vector (x, px, y, py, z, pz) plus charge and bookkeeping), so handing the reference
 * dense vectors costs a de-interleaving pass over the whole bunch before and after every step.
 * The *_strided entry points read and write such records in place: every array is a base pointer
 * plus an ELEMENT stride (in units of the pdt element, not bytes), particle i of array a lives at
 * a[i * stride_a].  All strides 1 = the dense calls above (same kernels, same results).
 *   x, y, z     >= 1
 *   q           >= 0; 0 = every particle carries the charge q[0] (equal-weight macro-particles:
 *               saves reading Np charges)
 *   ex, ey, ez  >= 1 (interpolated field, or the momenta updated by the fused kick); output
 *               elements of different particles must not alias.
 * Example, Bmad-style phase-space records `double vec[6]` in an array `v` of Np records, charges
 * in a dense vector: x = v, y = v + 2, z = v + 4 with strides 6; the kick updates px = v + 1,
 * py = v + 3, pz = v + 5 with strides 6 -- the bunch is never copied.
 * Arithmetic, clamping and error codes are those of the dense calls. */
typedef struct scb_particle_strides {
    int64_t x, y, z, q;        /* inputs (q is ignored by the calls that take no charges)          */
    int64_t ex, ey, ez;        /* outputs (ignored by scb_deposit_strided / scb_bounds_strided)    */
    int64_t reserved;
} scb_particle_strides;

SCB_API int scb_deposit_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                                const void* q, const scb_particle_strides* st, int pdt, void* rho, int mdt,
                                const int64_t n[3], const double min_bounds[3], const double delta[3],
                                int clear);
SCB_API int scb_interpolate_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                                    const scb_particle_strides* st, int pdt, const void* efield, int mdt,
                                    const int64_t n[3], const double min_bounds[3], const double delta[3],
                                    void* ex, void* ey, void* ez);
SCB_API int scb_interpolate_kick_strided(scb_handle* h, int64_t np, const void* x, const void* y,
                                         const void* z, const scb_particle_strides* st, int pdt,
                                         const void* efield, int mdt, const int64_t n[3],
                                         const double min_bounds[3], const double delta[3], void* px,
                                         void* py, void* pz, double coef_xy, double coef_z);
SCB_API int scb_bounds_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                               const scb_particle_strides* st, int pdt, double out_min[3],
                               double out_max[3]);
SCB_API int scb_step_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                             const void* q, const scb_particle_strides* st, int pdt, void* rho,
                             void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                             const double max_bounds[3], const double delta[3], double gamma,
                             int at_cathode, void* ex, void* ey, void* ez);

/* ---- extension: bunches kept ordered by cell ---------------------------------------------------- */
/* The reference takes particles in whatever order the caller holds them (src/deposition.jl:218-247,
 * src/interpolation.jl:100-128); with a random order every corner update / corner read is its own 32-byte L2
 * transaction, which bounds both particle passes far below the HBM roofline.  A tracking loop can keep its bunch
 * ordered by cell instead: particles move a fraction of a cell per step, so the order decays slowly and a re-sort every
 * K steps is enough (bench.py reports the break-even K).
 *
 * scb_sort_particles computes the permutation that orders the bunch by linear cell index ix + nx*(iy + ny*iz) (the
 * index arithmetic of deposit!/interpolate_field, clamped to the grid; stable, so particles of one cell keep their
 * relative order): perm_out[i] = index of the particle that comes i-th, np < 2^31 entries of uint32.
 * scb_permute applies it to up to 8 per-particle arrays of one element type in one pass: dst[f][i] = src[f][perm[i]]
 * (dst must not alias src) -- coordinates, charge, momenta, whatever the caller carries.
 * scb_set_particle_order tells the handle which kernels the particle passes use:
 *   SCB_ORDER_RANDOM  the sector-transaction-minimising kernels for unordered bunches (default);
 *   SCB_ORDER_CELL    run-accumulating kernels: every lane walks consecutive particles, keeps the current cell's eight
 *                     corner sums (deposit) / 24 field values (gather) in registers, warps combine their open runs
 *                     with a segmented shuffle reduction before touching memory.  Results are correct for ANY order
 *                     (a cell change merely ends a run); the speed depends on how ordered the bunch is.
 *   SCB_ORDER_CELL_TILE  SCB_ORDER_CELL with a shared-memory tile of rho nodes under the deposit (the CTA adds finished
 *                     runs to the tile with shared-memory atomics and flushes it with coalesced reductions).  Same
 *                     results; measured SLOWER on B200 (shared Float64 atomics are compare-and-swap loops, and an
 *                     ordered bunch makes a CTA's lanes collide on a handful of nodes): kept for comparison only.
 *   SCB_ORDER_AUTO    the handle decides: every eighth deposit (and whenever the bunch's arrays change) it samples the
 *                     order (the measurement of scb_particle_order_fraction) and uses the SCB_ORDER_CELL kernels when at
 *                     least 55 % of the sampled pairs share a cell or are x-neighbours, the SCB_ORDER_RANDOM kernels
 *                     otherwise; the gather follows the deposit.  For callers that cannot know how ordered their bunch
 *                     is -- the default kernels are SLOWER on an ordered bunch than on a random one (every lane of a warp
 *                     reduces into the same node).  The probe synchronises the stream: not usable under stream capture.
 * scb_particle_order_fraction samples neighbouring particle pairs and returns the fraction that share a cell or sit in
 * x-adjacent cells (about 1 for an ordered bunch, about 0 for a random one); synchronous. */
typedef enum scb_particle_order {
    SCB_ORDER_RANDOM = 0, SCB_ORDER_CELL = 1, SCB_ORDER_CELL_TILE = 2, SCB_ORDER_AUTO = 3
} scb_particle_order;
SCB_API int scb_sort_particles(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt,
                               int mdt, const int64_t n[3], const double min_bounds[3], const double delta[3],
                               uint32_t* perm_out);
SCB_API int scb_permute(scb_handle* h, int64_t np, const uint32_t* perm, int nfields, const void* const* src,
                        void* const* dst, int dt);
SCB_API int scb_set_particle_order(scb_handle* h, int order);
SCB_API int scb_particle_order_fraction(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                                        int pdt, int mdt, const int64_t n[3], const double min_bounds[3],
                                        const double delta[3], double* fraction_out);

/* ---- the same step with HOST particle buffers (pinned or pageable) ------------------------ */
/* Copies x,y,z,q host->device in chunks overlapped with deposition, solves, interpolates in
 * chunks overlapped with the device->host copy of ex,ey,ez.  rho/efield stay on the device
 * (pointers owned by the caller).  Synchronous on return. */
SCB_API int scb_step_host(scb_handle* h, int64_t np, const void* x_host, const void* y_host,
                  const void* z_host, const void* q_host, int pdt, void* rho, void* efield,
                  int mdt, const int64_t n[3], const double min_bounds[3],
                  const double max_bounds[3], const double delta[3], double gamma,
                  int at_cathode, void* ex_host, void* ey_host, void* ez_host);

/* Stream-pipelined form of scb_step_host for callers that push several independent bunches (or
 * repeat the step on fresh host data): returns as soon as the copies and kernels are queued.  Up to
 * two steps may be in flight -- they alternate between two device staging slots, so the host->device
 * copy of step k+1 runs while the device->host copy of step k is still draining (PCIe is full
 * duplex).  The host buffers of a step (inputs and outputs) must stay untouched, and the output
 * buffers of two consecutive steps must be distinct, until scb_step_host_wait returns.  rho/efield
 * hold the grids of the most recently queued step. */
SCB_API int scb_step_host_async(scb_handle* h, int64_t np, const void* x_host, const void* y_host,
                  const void* z_host, const void* q_host, int pdt, void* rho, void* efield,
                  int mdt, const int64_t n[3], const double min_bounds[3],
                  const double max_bounds[3], const double delta[3], double gamma,
                  int at_cathode, void* ex_host, void* ey_host, void* ez_host);
SCB_API int scb_step_host_wait(scb_handle* h);

/* ---- multi-GPU: particle-sharded step with a slab-decomposed solve (NCCL over NVLink) ------- */
/* One process and one handle per GPU.  scb_comm_unique_id fills 128 bytes on one rank; the caller
 * broadcasts them (torch.distributed / MPI) and every rank calls scb_comm_init.  NCCL is dlopen'ed
 * at that point (libnccl.so.2), so single-GPU users do not need it. */
SCB_API int scb_comm_unique_id(void* uid128);
SCB_API int scb_comm_init(scb_handle* h, int nranks, int rank, const void* uid128);
SCB_API int scb_comm_destroy(scb_handle* h);
/* Sum of the ranks' charge grids in place (ncclAllReduce on the library's communicator, handle's stream): the
 * "solve replicated" mode of a particle-sharded run -- every rank then calls scb_solve on the full grid. */
SCB_API int scb_allreduce_rho(scb_handle* h, void* rho, const int64_t n[3], int mdt);
/* rho_partial: this rank's un-reduced charge grid (full size).  Reduce-scatters it into z slabs,
 * runs the FFT passes slab-decomposed (all-to-all pencil transposes fused into the pass
 * addressing), and all-gathers the field so that every rank ends with the full efield.
 * Requires nz and the padded y length to be multiples of nranks; free space or cathode. */
SCB_API int scb_solve_sharded(scb_handle* h, const void* rho_partial, void* efield, int mdt,
                              const int64_t n[3], const double min_bounds[3],
                              const double max_bounds[3], const double delta[3], double gamma,
                              int at_cathode);

/* deposit + scb_solve_sharded + interpolation in one call.  rho_partial receives this rank's partial
 * charge grid, efield the complete field, ex/ey/ez the field at this rank's particles.
 * SCB_GATHER_OVERLAP=1 (experimental, slower as measured) overlaps the all-gather of the field with the
 * interpolation: the z slabs are broadcast one by one from the middle of the grid outwards, repacked as
 * they land, and the gather runs in passes over the cells whose planes have arrived. */
SCB_API int scb_step_sharded(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                     const void* q, int pdt, void* rho_partial, void* efield, int mdt,
                     const int64_t n[3], const double min_bounds[3], const double max_bounds[3],
                     const double delta[3], double gamma, int at_cathode, void* ex, void* ey, void* ez);

/* scb_step_host_async for a particle shard: every rank passes its own HOST shard and output buffers, the grids are the
 * rank's partial rho / the full efield as in scb_step_sharded.  slab_solve != 0: slab-decomposed solve (same divisibility
 * rule as scb_solve_sharded); 0: rho all-reduced, solve replicated.  All ranks must queue the same sequence of steps;
 * completion with scb_step_host_wait.  (The reference has no host-buffer or multi-device entry point; this is the
 * sharded form of the benchmark body benchmark/full_pipeline_benchmark.jl:26-30.) */
SCB_API int scb_step_host_sharded_async(scb_handle* h, int64_t np, const void* x_host, const void* y_host,
                  const void* z_host, const void* q_host, int pdt, void* rho_partial, void* efield,
                  int mdt, const int64_t n[3], const double min_bounds[3],
                  const double max_bounds[3], const double delta[3], double gamma,
                  int at_cathode, int slab_solve, void* ex_host, void* ey_host, void* ez_host);

/* ---- cache control ------------------------------------------------------------------------ */
SCB_API int scb_drop_green_cache(scb_handle* h);
/* bytes of device workspace currently owned by the handle */
SCB_API int64_t scb_workspace_bytes(const scb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SPACECHARGE_B200_H */
