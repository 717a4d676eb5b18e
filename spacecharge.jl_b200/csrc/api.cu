// api.cu -- C ABI (include/spacecharge_b200.h): handle, workspace, Green-spectrum cache and the
// orchestration of the pass kernels.  No CPU fallback anywhere: every entry point either
// launches the sm_100a kernels or returns an error code.
#include "../../include/spacecharge_b200.h"
#include "fft_passes.cuh"
#include "kernels.h"

#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace scb;

namespace {

constexpr double kCLight = 299792458.0;            // src/utils.jl:7
constexpr double kFPEI = kCLight * kCLight * 1e-7;  // src/utils.jl:8
constexpr int kMaxFftLen = 2048;
constexpr int kMaxGreenEntries = 4;

struct GreenKey {
    int n[3];
    double delta[3];
    double gamma;
    double offset[3];
    int mdt;
    int kind;  // 0: free-space compressed S, 1: image (corr along z, negated) full, 2: general full
    bool operator==(const GreenKey& o) const {
        for (int a = 0; a < 3; ++a)
            if (n[a] != o.n[a] || delta[a] != o.delta[a] || offset[a] != o.offset[a]) return false;
        return gamma == o.gamma && mdt == o.mdt && kind == o.kind;
    }
};

struct GreenEntry {
    GreenKey key;
    void* data = nullptr;
    size_t bytes = 0;
    size_t cap = 0;       // allocated size (>= bytes when the buffer was recycled)
    long long scomp = 0;
    int ncomp = 3;        // 3: Ex,Ey,Ez; 4: + potential (icomp 0) as component 3
    unsigned long long stamp = 0;
    // free space, Lz = 512: second copy of S with kz' fastest, St[((c*ninner + kx)*(Ly/2+1) + ky')*PZ + kz'], read by
    // the even/odd-bin z pass (one contiguous row per line); t_off = byte offset inside data, 0 = absent
    size_t t_off = 0;
    int PZ = 0;
};

}  // namespace

struct scb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    scb_options opt{};
    std::string err;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    std::map<std::pair<int, int>, void*> twiddles;  // (N, is_f64) -> device table of N roots
    std::vector<GreenEntry> green;
    std::vector<std::pair<void*, size_t>> green_pool;  // retired spectrum buffers, reused by the next build
    unsigned long long stamp = 0;
    unsigned long long* d_bounds = nullptr;
    void* packed = nullptr;       // node-major copy of efield for the gather (32 bytes per node)
    size_t packed_bytes = 0;
    void* tiles = nullptr;        // cell-tile accumulator of the deposit (4 mesh elements per node)
    size_t tiles_bytes = 0;
    int64_t launches = 0;
    // timing
    bool timing = false;
    cudaEvent_t ev[16]{};
    bool ev_ready = false;
    scb_timing last{};
    bool t_dep = false, t_solve = false, t_interp = false, t_green = false, t_pass = false, t_coll = false;
    // host-step staging
    // multi-GPU
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    void* slab = nullptr;          // reduce-scattered rho slab
    size_t slab_bytes = 0;
    // peer-memory path: every rank maps the other ranks' arenas (CUDA IPC) and the pass kernels store
    // straight into the destination rank's receive buffers
    int p2p = -1;                  // -1 undecided, 0 off (NCCL send/recv), 1 on
    unsigned long long arena_gen = 0, peer_gen = ~0ull;
    // workspace of the slab-decomposed solve.  Separate from `arena` on purpose: it is (re)allocated ONLY inside
    // run_solve_sharded, from arguments every rank shares, so all ranks decide alike whether the collective
    // re-exchange of IPC handles (exchange_arenas) has to run -- single-rank users of `arena` (scb_solve, scb_green,
    // the sort, the probes) can no longer make one rank skip or enter that collective alone.
    void* sh_arena = nullptr;
    size_t sh_arena_bytes = 0;
    unsigned long long sh_gen = 0;
    std::vector<void*> peer_arena;
    char* d_ipc = nullptr;         // nranks * 64 bytes of IPC handles + 2 ints (flag, barrier)
    // host-buffer steps: two staging slots so that the upload of step k+1 overlaps the download of step k
    cudaStream_t copy_stream = nullptr;   // host -> device
    cudaStream_t d2h_stream = nullptr;    // device -> host
    void* stage[2] = {nullptr, nullptr};
    size_t stage_bytes[2] = {0, 0};
    std::vector<cudaEvent_t> chunk_ev[2];
    cudaEvent_t slot_in_free[2] = {nullptr, nullptr};    // compute stream: last kernel that reads the slot's x,y,z,q
    cudaEvent_t slot_out_free[2] = {nullptr, nullptr};   // d2h stream: last copy that reads the slot's ex,ey,ez
    bool slot_used[2] = {false, false};
    int host_steps_in_flight = 0;
    // slab-decomposed step: field slabs broadcast on comm_stream, repacked on pack_stream, gathered on the main stream
    cudaStream_t comm_stream = nullptr, pack_stream = nullptr;
    cudaEvent_t ev_field = nullptr, ev_bcast[SCB_MAX_RANKS] = {}, ev_pack[SCB_MAX_RANKS] = {};
    // cold geometry inside a fused step: the Green spectrum is rebuilt on green_stream (high priority) while the
    // deposit runs on the main stream; the solve waits for ev_green_done
    cudaStream_t green_stream = nullptr;
    cudaEvent_t ev_green_start = nullptr, ev_green_done = nullptr;
    bool green_pending = false;
    // SCB_ORDER_AUTO: which kernel family the last order probe chose, and how many deposits it still holds for
    int auto_choice = SCB_ORDER_RANDOM, auto_hold = 0;
    const void* auto_ptr = nullptr;
    int64_t auto_np = -1;
    // z-chunked B2 -> B3 hand-over (run_solve): B2 chunks on chunk_stream, B3 chunks on the handle's stream
    cudaStream_t chunk_stream = nullptr;
    cudaEvent_t ev_chunk_fork = nullptr, ev_chunk_ready[2] = {nullptr, nullptr}, ev_chunk_free[2] = {nullptr, nullptr};
};

namespace {

void close_peers(scb_handle* h);  // defined with the multi-GPU code below

int fail(scb_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

int cuda_fail(scb_handle* h, cudaError_t e, const char* where) {
    return fail(h, SCB_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define SCB_CUDA(h, call)                                   \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(h, e__, #call); \
    } while (0)

#define SCB_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != SCB_OK) return rc__; \
    } while (0)

struct NcclApi {
    void* dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.dl) return true;
    void* dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!dl) return false;
#define SCB_SYM(field, name)                                              \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(dl, name);           \
    if (!g_nccl.field) return false;
    SCB_SYM(GetUniqueId, "ncclGetUniqueId")
    SCB_SYM(CommInitRank, "ncclCommInitRank")
    SCB_SYM(CommDestroy, "ncclCommDestroy")
    SCB_SYM(ReduceScatter, "ncclReduceScatter")
    SCB_SYM(AllGather, "ncclAllGather")
    SCB_SYM(AllReduce, "ncclAllReduce")
    SCB_SYM(Send, "ncclSend")
    SCB_SYM(Recv, "ncclRecv")
    SCB_SYM(GroupStart, "ncclGroupStart")
    SCB_SYM(GroupEnd, "ncclGroupEnd")
    SCB_SYM(Broadcast, "ncclBroadcast")
    SCB_SYM(GetErrorString, "ncclGetErrorString")
#undef SCB_SYM
    g_nccl.dl = dl;
    return true;
}

#define SCB_NCCL(h, call)                                                                         \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) return fail(h, SCB_ERR_COMM, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

int padded_len(int n) {
    int L = 8;
    while (L < 2 * n) L *= 2;
    return L;
}

bool valid_dt(int dt) { return dt == SCB_F32 || dt == SCB_F64; }
size_t dt_size(int dt) { return dt == SCB_F64 ? 8 : 4; }

int check_grid(scb_handle* h, const int64_t n[3]) {
    if (!n) return fail(h, SCB_ERR_INVALID_ARG, "grid size pointer is null");
    for (int a = 0; a < 3; ++a) {
        if (n[a] < 2) return fail(h, SCB_ERR_INVALID_ARG, "All elements of grid_size must be at least 2.");
        if (2 * n[a] > kMaxFftLen) return fail(h, SCB_ERR_UNSUPPORTED, "grid dimension larger than 1024 is not supported");
    }
    return SCB_OK;
}

int ensure_arena(scb_handle* h, size_t bytes) {
    if (bytes <= h->arena_bytes) return SCB_OK;
    if (h->arena) {
        SCB_CUDA(h, cudaStreamSynchronize(h->stream));
        SCB_CUDA(h, cudaFree(h->arena));
        h->arena = nullptr;
        h->arena_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&h->arena, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, SCB_ERR_ALLOC, "workspace allocation of " + std::to_string(bytes) + " bytes failed");
    }
    h->arena_bytes = bytes;
    h->arena_gen++;
    return SCB_OK;
}

int ensure_sh_arena(scb_handle* h, size_t bytes) {
    if (bytes <= h->sh_arena_bytes) return SCB_OK;
    if (h->sh_arena) {
        SCB_CUDA(h, cudaStreamSynchronize(h->stream));
        close_peers(h);   // the peers' mappings of the old buffer are re-made by exchange_arenas
        SCB_CUDA(h, cudaFree(h->sh_arena));
        h->sh_arena = nullptr;
        h->sh_arena_bytes = 0;
    }
    if (cudaMalloc(&h->sh_arena, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, SCB_ERR_ALLOC, "sharded-solve workspace allocation of " + std::to_string(bytes) + " bytes failed");
    }
    h->sh_arena_bytes = bytes;
    h->sh_gen++;
    return SCB_OK;
}

template <typename T>
int get_twiddles(scb_handle* h, int N, const cx_t<T>** out) {
    const int is64 = sizeof(T) == 8;
    auto it = h->twiddles.find({N, is64});
    if (it != h->twiddles.end()) {
        *out = static_cast<const cx_t<T>*>(it->second);
        return SCB_OK;
    }
    std::vector<cx_t<T>> host(N);
    for (int k = 0; k < N; ++k) {
        // exact octant symmetry is not needed; long double keeps the table correctly rounded
        const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)N;
        host[k].x = (T)cosl(a);
        host[k].y = (T)sinl(a);
    }
    void* d = nullptr;
    SCB_CUDA(h, cudaMalloc(&d, sizeof(cx_t<T>) * N));
    SCB_CUDA(h, cudaMemcpy(d, host.data(), sizeof(cx_t<T>) * N, cudaMemcpyHostToDevice));
    h->twiddles[{N, is64}] = d;
    *out = static_cast<const cx_t<T>*>(d);
    return SCB_OK;
}

int ensure_packed(scb_handle* h, size_t bytes) {
    if (bytes <= h->packed_bytes) return SCB_OK;
    if (h->packed) {
        SCB_CUDA(h, cudaStreamSynchronize(h->stream));
        SCB_CUDA(h, cudaFree(h->packed));
        h->packed = nullptr;
        h->packed_bytes = 0;
    }
    if (cudaMalloc(&h->packed, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, SCB_ERR_ALLOC, "packed-field allocation failed");
    }
    h->packed_bytes = bytes;
    return SCB_OK;
}

// Kernel family for this call.  SCB_ORDER_AUTO: every eighth deposit (or when the bunch's array changes) a sample of
// neighbouring particle pairs is located (k_order_probe, ~50 us + one stream synchronisation) and the ordered kernels are
// chosen when at least 55 % of the pairs share a cell or sit in x-adjacent cells -- measured crossover: at 72 % (a drift of
// 0.1 cell) the ordered kernels win (5.8 against 6.4 ms per step), at 36 % (0.3 cell) they lose (8.3 ms).  The gather
// follows the deposit's choice.  Not for use under stream capture (the probe synchronises).
int resolve_order(scb_handle* h, bool probe_now, int64_t np, const void* x, const void* y, const void* z, int pdt, int mdt,
                  const Geom3& g, int* order) {
    *order = h->opt.particle_order;
    if (*order != SCB_ORDER_AUTO) return SCB_OK;
    if (probe_now && np >= 4096 && (h->auto_hold <= 0 || h->auto_ptr != x || h->auto_np != np)) {
        SCB_CUDA(h, launch_order_probe(pdt, mdt, np, x, y, z, g, h->d_bounds, h->stream));
        unsigned long long host[2];
        SCB_CUDA(h, cudaMemcpyAsync(host, h->d_bounds, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
        SCB_CUDA(h, cudaStreamSynchronize(h->stream));
        h->launches += 1;
        const double frac = host[1] ? (double)host[0] / (double)host[1] : 0.0;
        h->auto_choice = frac >= 0.55 ? SCB_ORDER_CELL : SCB_ORDER_RANDOM;
        h->auto_hold = 8;
        h->auto_ptr = x;
        h->auto_np = np;
    } else if (probe_now) {
        if (np < 4096) h->auto_choice = SCB_ORDER_RANDOM;
        h->auto_hold -= 1;
    }
    *order = h->auto_choice;
    return SCB_OK;
}

// deposit: cell-tile accumulation when there are enough particles to pay for zeroing and folding the
// tiles (deposit_mode 0 = auto, 1 = one thread per particle, 2 = lane pairs, 3 = tiles)
int run_deposit(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q, int pdt,
                void* rho, int mdt, const Geom3& g, bool clear, bool cleared_already, const PLayout* lay = nullptr) {
    const long long ng = (long long)g.n[0] * g.n[1] * g.n[2];
    int order = SCB_ORDER_RANDOM;
    if (!lay) SCB_TRY(resolve_order(h, true, np, x, y, z, pdt, mdt, g, &order));
    if (order != SCB_ORDER_RANDOM && !lay) {
        if (clear && !cleared_already) SCB_CUDA(h, cudaMemsetAsync(rho, 0, (size_t)ng * dt_size(mdt), h->stream));
        SCB_CUDA(h, launch_deposit_runs(pdt, mdt, np, x, y, z, q, rho, g, h->stream, order == SCB_ORDER_CELL_TILE));
        if (np > 0) h->launches += 1;
        return SCB_OK;
    }
    int mode = h->opt.deposit_mode;
    if (mode == 0) mode = np >= ng ? 3 : 2;
    if (mode == 3) {
        const size_t need = (size_t)4 * ng * dt_size(mdt);
        if (h->tiles_bytes < need) {
            if (h->tiles) { SCB_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(h->tiles); h->tiles = nullptr; h->tiles_bytes = 0; }
            if (cudaMalloc(&h->tiles, need) != cudaSuccess) { (void)cudaGetLastError(); mode = 2; }
            else h->tiles_bytes = need;
        }
    }
    if (mode == 3) {
        // a cleared rho equals "overwrite"; an un-cleared one is accumulated into
        SCB_CUDA(h, launch_deposit_tiles(pdt, mdt, np, x, y, z, q, h->tiles, rho, g, clear ? 0 : 1, h->stream, lay));
        h->launches += 2;
        return SCB_OK;
    }
    if (clear && !cleared_already) SCB_CUDA(h, cudaMemsetAsync(rho, 0, (size_t)ng * dt_size(mdt), h->stream));
    SCB_CUDA(h, launch_deposit(pdt, mdt, np, x, y, z, q, rho, g, mode, h->stream, lay));
    if (np > 0) h->launches += 1;
    return SCB_OK;
}

// gather: repack the field node-major when there are enough particles to pay for the extra pass
int run_interpolate(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt, const void* efield,
                    int mdt, const Geom3& g, void* ex, void* ey, void* ez, bool* packed_ready, const Kick& kick = Kick(),
                    const PLayout* lay = nullptr) {
    const long long ng = (long long)g.n[0] * g.n[1] * g.n[2];
    int order = SCB_ORDER_RANDOM;
    if (!lay) SCB_TRY(resolve_order(h, false, np, x, y, z, pdt, mdt, g, &order));
    if (order != SCB_ORDER_RANDOM && !lay && !(packed_ready && *packed_ready)) {
        // SCB_CELL_GATHER=1 (tuning): one thread per particle straight from efield instead of the run-accumulating walk
        static const int direct = [] { const char* e = std::getenv("SCB_CELL_GATHER"); return e ? std::atoi(e) : 0; }();
        if (direct == 1) SCB_CUDA(h, launch_interpolate(pdt, mdt, np, x, y, z, efield, g, ex, ey, ez, h->stream, kick, nullptr));
        else SCB_CUDA(h, launch_interpolate_runs(pdt, mdt, np, x, y, z, efield, g, ex, ey, ez, h->stream, kick));
        h->launches += 1;
        return SCB_OK;
    }
    const bool use_packed = (packed_ready && *packed_ready) || np * 4 >= ng;
    if (!use_packed) {
        SCB_CUDA(h, launch_interpolate(pdt, mdt, np, x, y, z, efield, g, ex, ey, ez, h->stream, kick, lay));
        h->launches += 1;
        return SCB_OK;
    }
    if (!(packed_ready && *packed_ready)) {
        SCB_TRY(ensure_packed(h, (size_t)ng * packed_bytes_per_node(mdt)));
        SCB_CUDA(h, launch_pack_efield(mdt, efield, h->packed, g, h->stream));
        h->launches += 1;
        if (packed_ready) *packed_ready = true;
    }
    SCB_CUDA(h, launch_interpolate_packed(pdt, mdt, np, x, y, z, h->packed, g, ex, ey, ez, h->stream, kick, lay));
    h->launches += 1;
    return SCB_OK;
}

// Stage boundaries: CUDA events for scb_get_timing (when enabled) and NVTX ranges for profilers (always; SURVEY.md
// section 5, tracing).  Ranges nest as  scb:solve > { scb:green_build, scb:F1 ... scb:B3 }.
void tick(scb_handle* h, int i) {
    switch (i) {
        case 0: nvtxRangePushA("scb:deposit"); break;
        case 2: nvtxRangePushA("scb:solve"); break;
        case 4: nvtxRangePushA("scb:interpolate"); break;
        case 6: nvtxRangePushA("scb:green_build"); break;
        case 8: nvtxRangePushA(h->nranks > 1 && h->comm ? "scb:F1 x r2c (+ reduce-scatter of rho when sharded)" : "scb:F1 x r2c"); break;
        case 9: nvtxRangePop(); nvtxRangePushA("scb:F2 y forward (+ pencil transpose when sharded)"); break;
        case 10: nvtxRangePop(); nvtxRangePushA("scb:Z fused z pass: fft, Green multiply, ifft x3"); break;
        case 11: nvtxRangePop(); nvtxRangePushA("scb:B2 y inverse"); break;
        case 12: nvtxRangePop(); nvtxRangePushA("scb:B3 x c2r (+ all-gather of E when sharded)"); break;
        case 14: nvtxMarkA("scb:reduce-scatter queued"); break;
        case 15: nvtxMarkA("scb:all-gather next"); break;
        case 1: case 3: case 5: case 7: case 13: nvtxRangePop(); break;
        default: break;
    }
    if (h->timing && h->ev_ready) cudaEventRecord(h->ev[i], h->stream);
}

struct Plan {
    int n[3];
    int L[3];
    int PX, ninner;
};

Plan make_plan(const int64_t n[3]) {
    Plan p;
    for (int a = 0; a < 3; ++a) {
        p.n[a] = (int)n[a];
        p.L[a] = padded_len((int)n[a]);
    }
    p.ninner = p.L[0] / 2 + 1;
    p.PX = (p.ninner + 7) / 8 * 8;
    return p;
}

// ---- Green spectrum: build (cold) and cache ------------------------------------------------
// the even/odd-bin z pass (k_z_eo) exists for a padded z length of 512 (129..256 grid points along z)
bool z_eo_eligible(const Plan& pl) { return pl.L[2] == 512; }
// SCB_GREEN_REAL=0 selects the complex passes + compress kernel for the free-space build (A/B timing, debugging)
bool real_sym_disabled() {
    static const bool off = [] { const char* e = getenv("SCB_GREEN_REAL"); return e && e[0] == '0'; }();
    return off;
}

int build_green(scb_handle* h, const Plan& pl, const GreenKey& key, GreenEntry& ent, int ncomp) {
    IgfGeom g{};
    for (int a = 0; a < 3; ++a) {
        g.n[a] = pl.n[a];
        g.L[a] = pl.L[a];
        g.isize[a] = 2 * pl.n[a];
        g.delta[a] = key.delta[a];
        g.offset[a] = key.offset[a];
        g.sym[a] = key.offset[a] == 0.0 ? 1 : 0;
        g.corr[a] = 0;
    }
    g.gamma = key.gamma;
    double sign_all = 1.0;
    if (key.kind == 1) {  // image charge: correlation along z, negated (src/solvers/free_space.jl:34)
        g.corr[2] = 1;
        g.sym[2] = 0;
        sign_all = -1.0;
    }
    for (int a = 0; a < 3; ++a) {
        g.i0[a] = g.sym[a] ? pl.n[a] : 1;
        g.cnt[a] = g.sym[a] ? pl.n[a] + 1 : 2 * pl.n[a];
    }
    const size_t nP = (size_t)g.cnt[0] * g.cnt[1] * g.cnt[2];
    const size_t nS = (size_t)pl.PX * pl.L[1] * pl.L[2];
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    // free space and image charge: only ky <= Ly/2 and kz <= Lz/2 are kept (the y pass prunes its output,
    // the z pass runs on half of the lines and prunes too).  Free space: the spectrum is i*S with S real and
    // (anti)symmetric.  Image charge: h is real, even/odd in x and y with parities p_x, p_y, general in z, so
    // H(kx, Ly-ky, kz) = p_y H(kx,ky,kz) and H(kx, ky, Lz-kz) = p_x p_y conj(H(kx,ky,kz)).
    const int Lyh1 = pl.L[1] / 2 + 1, Lzh1 = pl.L[2] / 2 + 1;
    const bool prune = key.kind == 0 || key.kind == 1;
    const size_t nY2 = nS;   // second full-size buffer: the passes ping-pong between the two
    const size_t need = 2 * al(nP * 8) + al(nS * 16) + al(nY2 * 16);
    SCB_TRY(ensure_arena(h, need));
    char* base = static_cast<char*>(h->arena);
    double* P = reinterpret_cast<double*>(base);
    double* Dd = reinterpret_cast<double*>(base + al(nP * 8));
    double2* spec = reinterpret_cast<double2*>(base + 2 * al(nP * 8));
    double2* Y2 = reinterpret_cast<double2*>(base + 2 * al(nP * 8) + al(nS * 16));

    const bool f64 = key.mdt == SCB_F64;
    size_t per_comp;  // elements
    if (prune) per_comp = (size_t)pl.PX * Lyh1 * Lzh1;
    else per_comp = nS;
    const size_t elem = (key.kind == 0 ? 1 : 2) * dt_size(key.mdt);
    ent.bytes = (size_t)ncomp * per_comp * elem;
    ent.scomp = (long long)per_comp;
    ent.ncomp = ncomp;
    ent.t_off = 0;
    ent.PZ = 0;
    if ((key.kind == 0 || key.kind == 1) && z_eo_eligible(pl)) {   // kz'-fastest copy for the even/odd-bin z pass
        ent.PZ = (Lzh1 + 7) / 8 * 8;
        ent.t_off = al(ent.bytes);
        ent.bytes = ent.t_off + (size_t)ncomp * pl.ninner * Lyh1 * ent.PZ * elem;
    }
    // reuse a retired buffer when one is large enough: cudaFree/cudaMalloc of ~0.5 GB blocks costs
    // anything from 1 to 400 ms on the host and would dominate a re-mesh-every-step workload
    ent.data = nullptr;
    for (size_t i = 0; i < h->green_pool.size(); ++i) {
        if (h->green_pool[i].second >= ent.bytes && h->green_pool[i].second <= 2 * ent.bytes) {
            ent.data = h->green_pool[i].first;
            ent.cap = h->green_pool[i].second;
            h->green_pool.erase(h->green_pool.begin() + i);
            break;
        }
    }
    if (!ent.data) {
        cudaError_t e = cudaMalloc(&ent.data, ent.bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ent.data = nullptr;
            return fail(h, SCB_ERR_ALLOC, "Green-spectrum allocation failed");
        }
        ent.cap = ent.bytes;
    }

    const double2 *twx, *twy, *twz;
    SCB_TRY(get_twiddles<double>(h, pl.L[0], &twx));
    SCB_TRY(get_twiddles<double>(h, pl.L[1], &twy));
    SCB_TRY(get_twiddles<double>(h, pl.L[2], &twz));

    // Symmetric axes (zero offset, convolution placement): the padded IGF is even, or odd along the
    // component's own axis, about index 0, so only rows Y <= Ly/2 (Z <= Lz/2) are transformed along x
    // and the y (z) pass reads the missing half of each line mirrored with the parity sign.
    const bool foldY = g.sym[1] && !g.corr[1];
    const bool foldZ = g.sym[2] && !g.corr[2];
    const int rowsY = foldY ? Lyh1 : pl.L[1];
    const int rowsZ = foldZ ? Lzh1 : pl.L[2];
    const bool real_sym = key.kind == 0 && g.sym[0] && g.sym[1] && g.sym[2] && !real_sym_disabled();
    for (int c = 0; c < ncomp; ++c) {
        // components 0..2 = reference icomp 1..3 (src/green_functions.jl:90-98); component 3 = potential (icomp 0)
        const int icomp = c < 3 ? c + 1 : 0;
        SCB_CUDA(h, launch_green_point(P, g, icomp, h->stream));
        SCB_CUDA(h, launch_green_diff(Dd, P, g, h->stream));
        XParams<double> xp{};
        // the padded real array is generated inside the x pass (never written to memory)
        xp.gen.D = Dd;
        for (int a = 0; a < 3; ++a) {
            xp.gen.n[a] = g.n[a]; xp.gen.L[a] = g.L[a]; xp.gen.sym[a] = g.sym[a]; xp.gen.corr[a] = g.corr[a];
            xp.gen.dcnt[a] = g.cnt[a] - 1;
        }
        xp.gen.icomp = icomp;
        xp.gen.ly_lines = rowsY;
        xp.gen.sign_all = sign_all;
        xp.in = nullptr;
        xp.out = spec;                                   // [kx][Y < rowsY][Z < rowsZ]
        xp.tw = twx;
        xp.nlines = (long long)rowsY * rowsZ;
        xp.real_sline = pl.L[0];
        xp.n_real = pl.L[0];
        xp.PX = pl.PX;
        xp.scale = 1.0;
        char* dst = static_cast<char*>(ent.data) + (size_t)c * per_comp * elem;
        if (real_sym) {
            // Free space: the padded IGF is real and even / odd about 0 along every axis, so the spectrum stays
            // "real up to a power of i" through all three passes.  The x pass stores only that real part; the y and
            // z passes then see PAIRS of adjacent kx lines as one complex line (a + i b): half the lines, half the
            // bytes, and the last pass writes the cached real spectrum S directly (Float64) -- no compress kernel.
            const int PXc = pl.PX / 2;
            xp.real_out = (c == 0) ? 2 : 1;                  // odd along x only for E_x
            SCB_CUDA(h, launch_x_r2c<double>(pl.L[0], xp, 1, h->stream));
            LinesParams<double> yp{};
            yp.in = spec;                                    // pairs [kx/2][Y < rowsY][Z < rowsZ]
            yp.out = Y2;                                     // pairs [kx/2][ky < Lyh1][Z < rowsZ]
            yp.tw = twy;
            yp.n_in = pl.L[1];
            yp.n_out = Lyh1;
            yp.ninner = PXc;
            yp.in_sline = yp.out_sline = PXc;
            yp.in_souter = (long long)PXc * rowsY;
            yp.out_souter = (long long)PXc * Lyh1;
            yp.in_fold = 1;
            yp.fold_sign = (c == 1) ? -1.0 : 1.0;
            yp.out_rot = (c == 1) ? 1 : 0;
            yp.scale = 1.0;
            SCB_CUDA(h, launch_lines<double>(pl.L[1], -1, yp, rowsZ, 1, h->stream));
            LinesParams<double> zp{};
            zp.in = Y2;
            zp.out = f64 ? reinterpret_cast<double2*>(dst) : spec;   // pairs [kx/2][ky < Lyh1][kz < Lzh1]
            zp.tw = twz;
            zp.n_in = pl.L[2];
            zp.n_out = Lzh1;
            zp.ninner = PXc;
            zp.in_sline = zp.out_sline = (long long)PXc * Lyh1;
            zp.in_souter = zp.out_souter = PXc;
            zp.in_fold = 1;
            zp.fold_sign = (c == 2) ? -1.0 : 1.0;
            zp.out_rot = (c == 2) ? 1 : 0;
            zp.scale = 1.0;
            SCB_CUDA(h, launch_lines<double>(pl.L[2], -1, zp, Lyh1, 1, h->stream));
            h->launches += 5;
            if (!f64) {
                SCB_CUDA(h, launch_green_real_to_f32(dst, reinterpret_cast<const double*>(spec), (long long)per_comp, h->stream));
                h->launches += 1;
            }
            continue;
        }
        SCB_CUDA(h, launch_x_r2c<double>(pl.L[0], xp, 1, h->stream));
        const int ly_out = prune ? Lyh1 : pl.L[1];   // ky kept after the y pass
        const int lz_out = prune ? Lzh1 : pl.L[2];
        const bool y_in_place = !foldY && !prune;
        double2* ydst = y_in_place ? spec : Y2;          // [kx][ky < ly_out][Z < rowsZ]
        LinesParams<double> yp{};
        yp.in = spec;
        yp.out = ydst;
        yp.tw = twy;
        yp.n_in = pl.L[1];
        yp.n_out = ly_out;
        yp.ninner = pl.ninner;
        yp.in_sline = yp.out_sline = pl.PX;
        yp.in_souter = (long long)pl.PX * rowsY;
        yp.out_souter = (long long)pl.PX * ly_out;
        yp.in_fold = foldY ? 1 : 0;
        yp.fold_sign = (c == 1) ? -1.0 : 1.0;
        yp.scale = 1.0;
        SCB_CUDA(h, launch_lines<double>(pl.L[1], -1, yp, rowsZ, 1, h->stream));
        const bool z_in_place = y_in_place && !foldZ;    // same strides in and out
        double2* zdst = z_in_place ? spec : (ydst == spec ? Y2 : spec);   // [kx][ky < ly_out][kz < lz_out]
        LinesParams<double> zp{};
        zp.in = ydst;
        zp.out = zdst;
        zp.tw = twz;
        zp.n_in = pl.L[2];
        zp.n_out = lz_out;
        zp.ninner = pl.ninner;
        zp.in_sline = zp.out_sline = (long long)pl.PX * ly_out;
        zp.in_souter = zp.out_souter = pl.PX;
        zp.in_fold = foldZ ? 1 : 0;
        zp.fold_sign = (c == 2) ? -1.0 : 1.0;
        zp.scale = 1.0;
        SCB_CUDA(h, launch_lines<double>(pl.L[2], -1, zp, ly_out, 1, h->stream));
        const double2* final_spec = zdst;
        if (key.kind == 0)
            SCB_CUDA(h, launch_green_compress_free(dst, f64, final_spec, pl.ninner, pl.PX, Lyh1, Lzh1, icomp == 0, h->stream));
        else
            SCB_CUDA(h, launch_green_convert_full(dst, f64, final_spec, pl.ninner, pl.PX, (long long)per_comp, h->stream));
        h->launches += 7;
    }
    if (ent.t_off) {
        for (int c = 0; c < ncomp; ++c) {
            const char* src = static_cast<const char*>(ent.data) + (size_t)c * per_comp * elem;
            char* dstT = static_cast<char*>(ent.data) + ent.t_off + (size_t)c * pl.ninner * Lyh1 * ent.PZ * elem;
            SCB_CUDA(h, launch_green_transpose(dstT, src, f64, pl.ninner, pl.PX, Lyh1, Lzh1, ent.PZ, h->stream, key.kind == 1));
        }
        h->launches += ncomp;
    }
    return SCB_OK;
}

// stream-ordered reuse is safe: every consumer of a retired buffer was enqueued on h->stream before
// the next build's kernels
void retire_green(scb_handle* h, GreenEntry& e) {
    if (!e.data) return;
    if (h->green_pool.size() < 4) h->green_pool.push_back({e.data, e.cap});
    else cudaFree(e.data);
    e.data = nullptr;
}

void free_green(scb_handle* h, bool release_pool) {
    for (auto& e : h->green) retire_green(h, e);
    h->green.clear();
    if (release_pool) {
        for (auto& b : h->green_pool) cudaFree(b.first);
        h->green_pool.clear();
    }
}

int get_green(scb_handle* h, const Plan& pl, const GreenKey& key, const GreenEntry** out, int ncomp = 3) {
    if (h->opt.green_cache) {
        for (size_t i = 0; i < h->green.size(); ++i) {
            GreenEntry& e = h->green[i];
            if (e.key == key) {
                if (e.ncomp >= ncomp) {
                    e.stamp = ++h->stamp;
                    *out = &e;
                    return SCB_OK;
                }
                // cached without the potential component: rebuild with it
                retire_green(h, e);
                h->green.erase(h->green.begin() + i);
                break;
            }
        }
    } else {
        // reference behaviour: the Green function is recomputed on every solve
        for (size_t i = 0; i < h->green.size();) {
            if (h->green[i].key.kind == key.kind) {
                retire_green(h, h->green[i]);
                h->green.erase(h->green.begin() + i);
            } else {
                ++i;
            }
        }
    }
    if ((int)h->green.size() >= kMaxGreenEntries) {
        size_t victim = 0;
        for (size_t i = 1; i < h->green.size(); ++i)
            if (h->green[i].stamp < h->green[victim].stamp) victim = i;
        retire_green(h, h->green[victim]);
        h->green.erase(h->green.begin() + victim);
    }
    GreenEntry ent;
    ent.key = key;
    ent.stamp = ++h->stamp;
    tick(h, 6);
    int rc = build_green(h, pl, key, ent, ncomp);
    tick(h, 7);
    h->t_green = true;
    if (rc != SCB_OK) {
        retire_green(h, ent);
        return rc;
    }
    h->green.push_back(ent);
    *out = &h->green.back();
    return SCB_OK;
}

GreenKey make_key(const Plan& pl, const double delta[3], double gamma, const double offset[3], int mdt, int kind) {
    GreenKey k;
    std::memset(&k, 0, sizeof(k));
    for (int a = 0; a < 3; ++a) {
        k.n[a] = pl.n[a];
        k.delta[a] = delta[a];
        k.offset[a] = offset[a] == 0.0 ? 0.0 : offset[a];  // fold -0.0
    }
    k.gamma = gamma;
    k.mdt = mdt;
    k.kind = kind;
    return k;
}

// ---- tensor maps for the TMA variant of the fused z pass ---------------------------------------
typedef CUresult (*scb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

scb_encode_tiled_fn tensor_map_encoder() {
    static scb_encode_tiled_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<scb_encode_tiled_fn>(p);
    }();
    return fn;
}

// dims / box in elements (innermost first), strides in bytes for dims 1..rank-1
bool make_tensor_map(CUtensorMap* m, bool f64, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                     const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
    scb_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    static const int promo = [] { const char* e = std::getenv("SCB_TMA_L2"); return e ? std::atoi(e) : 0; }();   // tuning knob
    const CUtensorMapL2promotion l2 = promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                      : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    return enc(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
               const_cast<void*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
               l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool z_eo_enabled() {
    static const bool on = [] { const char* e = std::getenv("SCB_Z_EO"); return !e || std::atoi(e) != 0; }();
    return on;
}

bool z_tma_enabled() {
    static const bool on = [] { const char* e = std::getenv("SCB_Z_TMA"); return !e || std::atoi(e) != 0; }();
    return on;
}

// planes per chunk of the B2 -> B3 hand-over through the L2 (0 = off: one launch each, D through HBM)
int zchunk_planes(const Plan& /*pl*/) {
    static const int env = [] { const char* e = std::getenv("SCB_ZCHUNK"); return e ? std::atoi(e) : -1; }();
    if (env >= 0) return env;
    return 0;
}

int ensure_chunk_streams(scb_handle* h) {
    if (h->chunk_stream) return SCB_OK;
    SCB_CUDA(h, cudaStreamCreateWithFlags(&h->chunk_stream, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&h->ev_chunk_fork, &h->ev_chunk_ready[0], &h->ev_chunk_ready[1], &h->ev_chunk_free[0], &h->ev_chunk_free[1]})
        SCB_CUDA(h, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return SCB_OK;
}

// The fused z pass on data of pitch p.PX (the full PX on one GPU, PX / ranks in the kx-slab solve): the even/odd-bin
// kernel when it applies (free space, padded z length 512), else the TMA kernel (free space, nz <= 256), else k_z_fused
// (cathode image, general offset, long z).  `p` arrives with everything but the tensor maps' business filled in.
template <typename T>
int launch_z_pass(scb_handle* h, const Plan& pl, ZParams<T>& p, int mode, const GreenEntry* gfree, const GreenEntry* gaux,
                  const cx_t<T>* B, cx_t<T>* Cc, size_t comp_stride, int nc) {
    const int kind = mode == 0 ? GREEN_FREE : mode == 1 ? GREEN_CATHODE : GREEN_FULL;
    const bool f64 = sizeof(T) == 8;
    const cuuint64_t s = sizeof(T), PX = p.PX, L1 = pl.L[1], nz = pl.n[2];
    const bool eo_free = mode == 0, eo_cath = mode == 1 && gaux && gaux->t_off && gaux->PZ == ZEoLayout<T>::PZ;
    if ((eo_free || eo_cath) && z_eo_enabled() && z_eo_eligible(pl) && gfree->t_off && gfree->PZ == ZEoLayout<T>::PZ) {
        // even/odd-bin variant: one warp per line, 256-point transforms with a single exchange (see k_z_eo)
        const cuuint32_t TX = ZEoLayout<T>::TX;
        const int rowb = ZEoLayout<T>::ROWB;
        const CUtensorMapSwizzle swz = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        CUtensorMap mB, mC;
        const cuuint64_t dB[3] = {2 * PX, L1, nz}, sB[2] = {2 * PX * s, 2 * PX * L1 * s};
        const cuuint32_t bB[3] = {2 * TX, 1, (cuuint32_t)nz};
        const cuuint64_t dC[4] = {2 * PX, L1, nz, (cuuint64_t)nc}, sC[3] = {2 * PX * s, 2 * PX * L1 * s, (cuuint64_t)comp_stride * 2 * s};
        const cuuint32_t bC[4] = {2 * TX, 1, (cuuint32_t)nz, 1};
        if (make_tensor_map(&mB, f64, 3, B, dB, sB, bB, swz) && make_tensor_map(&mC, f64, 4, Cc, dC, sC, bC, swz)) {
            p.St = reinterpret_cast<const T*>(static_cast<const char*>(gfree->data) + gfree->t_off);
            p.PZ = gfree->PZ;
            if (eo_cath) p.Ht = reinterpret_cast<const cx_t<T>*>(static_cast<const char*>(gaux->data) + gaux->t_off);
            SCB_CUDA(h, launch_z_eo<T>(p, mB, mC, h->stream, eo_cath));
            return SCB_OK;
        }
    }
    // (kx-slab solve: the spectrum tile of k_z_tma starts at this rank's global kx0; a start that is not 16-byte aligned
    // made the bulk tensor load trap -- illegal instruction on the odd ranks of a 4-rank Float32 run -- so those ranks
    // take k_z_fused)
    if (mode == 0 && z_tma_enabled() && pl.n[2] <= 256 && pl.L[2] >= 16 && pl.L[2] <= 512 &&
        ((size_t)p.kx0 * sizeof(T)) % 16 == 0) {
        // all global traffic of the pass through the TMA unit (see k_z_tma)
        const cuuint32_t TX = (cuuint32_t)tz_for(pl.L[2]);
        const cuuint64_t PXg = pl.PX, Lyh1 = pl.L[1] / 2 + 1, Lzh1 = pl.L[2] / 2 + 1;
        CUtensorMap mB, mC, mS;
        const cuuint64_t dB[3] = {2 * PX, L1, nz}, sB[2] = {2 * PX * s, 2 * PX * L1 * s};
        const cuuint32_t bB[3] = {2 * TX, 1, (cuuint32_t)nz};
        const cuuint64_t dC[4] = {2 * PX, L1, nz, (cuuint64_t)nc}, sC[3] = {2 * PX * s, 2 * PX * L1 * s, (cuuint64_t)comp_stride * 2 * s};
        const cuuint32_t bC[4] = {2 * TX, 1, (cuuint32_t)nz, 1};
        const cuuint64_t dS[4] = {PXg, Lyh1, Lzh1, (cuuint64_t)gfree->ncomp},
                         sS[3] = {PXg * s, PXg * Lyh1 * s, (cuuint64_t)gfree->scomp * s};
        const cuuint32_t bS[4] = {TX, 1, (cuuint32_t)z_tma_srows<T>(pl.L[2]), 1};
        if (make_tensor_map(&mB, f64, 3, B, dB, sB, bB) && make_tensor_map(&mC, f64, 4, Cc, dC, sC, bC) &&
            make_tensor_map(&mS, f64, 4, gfree->data, dS, sS, bS)) {
            cudaError_t e = launch_z_tma<T>(pl.L[2], p, mB, mC, mS, h->stream);
            if (e == cudaSuccess) return SCB_OK;
            if (e != cudaErrorNotSupported) return cuda_fail(h, e, "launch_z_tma");
            (void)cudaGetLastError();
        }
    }
    SCB_CUDA(h, launch_z_fused<T>(pl.L[2], kind, p, h->stream));
    return SCB_OK;
}

// ---- the convolution -----------------------------------------------------------------------
// mode 0: free space; mode 1: free space + cathode image (offset_z given); mode 2: general offset
// phi (optional): scalar potential as a fourth component through the same passes (extension)
template <typename T>
int run_solve(scb_handle* h, const T* rho, T* efield, const Plan& pl, const double delta[3], double gamma,
              int mode, const double offset[3], T* phi = nullptr) {
    using C = cx_t<T>;
    const int nc = phi ? 4 : 3;
    const int mdt = sizeof(T) == 8 ? SCB_F64 : SCB_F32;
    const double zero3[3] = {0, 0, 0};
    const GreenEntry* gfree = nullptr;
    const GreenEntry* gaux = nullptr;
    h->t_green = false;
    // the builds use the arena as scratch, so they must come before the passes touch it
    if (mode == 0 || mode == 1) SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, zero3, mdt, 0), &gfree, nc));
    if (mode == 1) {
        SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, offset, mdt, 1), &gaux, nc));
        // get_green may have reallocated the vector
        for (auto& e : h->green)
            if (e.key == make_key(pl, delta, gamma, zero3, mdt, 0)) gfree = &e;
    }
    if (mode == 2) SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, offset, mdt, 2), &gaux, nc));
    if (h->green_pending) {   // spectrum built on green_stream by prefetch_green: join before the passes
        SCB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_green_done, 0));
        h->green_pending = false;
    }

    const size_t szA = (size_t)pl.PX * pl.n[1] * pl.n[2];
    const size_t szB = (size_t)pl.PX * pl.L[1] * pl.n[2];
    SCB_TRY(ensure_arena(h, ((1 + nc) * szA + (1 + nc) * szB) * sizeof(C)));
    C* A = static_cast<C*>(h->arena);
    C* B = A + szA;
    C* Cc = B + szB;
    C* D = Cc + nc * szB;

    const C *twx, *twy, *twz;
    SCB_TRY(get_twiddles<T>(h, pl.L[0], &twx));
    SCB_TRY(get_twiddles<T>(h, pl.L[1], &twy));
    SCB_TRY(get_twiddles<T>(h, pl.L[2], &twz));

    tick(h, 8);
    {  // F1
        XParams<T> p{};
        p.in = rho;
        p.out = A;
        p.tw = twx;
        p.nlines = (long long)pl.n[1] * pl.n[2];
        p.real_sline = pl.n[0];
        p.n_real = pl.n[0];
        p.PX = pl.PX;
        p.scale = (T)1;
        SCB_CUDA(h, launch_x_r2c<T>(pl.L[0], p, 1, h->stream));
    }
    tick(h, 9);
    {  // F2
        LinesParams<T> p{};
        p.in = A;
        p.out = B;
        p.tw = twy;
        p.n_in = pl.n[1];
        p.n_out = pl.L[1];
        p.ninner = pl.ninner;
        p.in_sline = pl.PX;
        p.in_souter = (long long)pl.PX * pl.n[1];
        p.out_sline = pl.PX;
        p.out_souter = (long long)pl.PX * pl.L[1];
        p.scale = (T)1;
        SCB_CUDA(h, launch_lines<T>(pl.L[1], -1, p, pl.n[2], 1, h->stream));
    }
    tick(h, 10);
    {  // Z
        ZParams<T> p{};
        p.in = B;
        p.out = Cc;
        p.tw = twz;
        p.out_scomp = (long long)szB;
        p.nz = pl.n[2];
        p.ncomp = nc;
        p.ninner = pl.ninner;
        p.PX = pl.PX;
        p.Ly = pl.L[1];
        p.Lyg = pl.L[1];
        p.ky0 = 0;
        if (gfree) {
            p.S = static_cast<const T*>(gfree->data);
            p.S_scomp = gfree->scomp;
        }
        if (gaux) {
            p.H = static_cast<const C*>(gaux->data);
            p.H_scomp = gaux->scomp;
        }
        SCB_TRY(launch_z_pass<T>(h, pl, p, mode, gfree, gaux, B, Cc, szB, nc));
    }
    tick(h, 11);
    // B2 + B3.  Plain form: B2 writes the y-pruned intermediate D (3A bytes) to HBM and B3 reads it back.  Chunked form
    // (zchunk planes at a time, SCB_ZCHUNK): B2 of chunk k writes a small ring slot that B3 of chunk k consumes while it
    // is still in the L2, and that the next chunk but one overwrites before it is ever written back -- D never touches
    // HBM (algorithmic traffic 3B + 3*Ng*s instead of 3B + 6A + 3*Ng*s).  B2 runs on a side stream, B3 on the handle's
    // stream, two ring slots, so the tail of one chunk's launch overlaps the head of the next.
    LinesParams<T> pb2{};
    pb2.tw = twy;
    pb2.n_in = pl.L[1];
    pb2.n_out = pl.n[1];
    pb2.ninner = pl.ninner;
    pb2.in_sline = pl.PX;
    pb2.in_souter = (long long)pl.PX * pl.L[1];
    pb2.out_sline = pl.PX;
    pb2.out_souter = (long long)pl.PX * pl.n[1];
    pb2.in_scomp = (long long)szB;
    pb2.scale = (T)1;
    XParams<T> pb3{};
    pb3.tw = twx;
    pb3.real_sline = pl.n[0];
    pb3.n_real = pl.n[0];
    pb3.PX = pl.PX;
    pb3.real_scomp = (long long)pl.n[0] * pl.n[1] * pl.n[2];
    // factr = T(FPEI) (src/solvers/free_space.jl:75) times the inverse-FFT 1/M (:95)
    pb3.scale = (T)(kFPEI / ((double)pl.L[0] * pl.L[1] * pl.L[2]));
    const int zc = zchunk_planes(pl);
    if (zc > 0 && 2 * zc <= pl.n[2] && ensure_chunk_streams(h) == SCB_OK) {
        const size_t plane = (size_t)pl.PX * pl.n[1];          // complex elements of one z plane of D
        const size_t slot = (size_t)nc * plane * zc;             // one ring slot: nc components x zc planes
        // the ring lives at the start of D (D itself is not used in this form)
        SCB_CUDA(h, cudaEventRecord(h->ev_chunk_fork, h->stream));
        SCB_CUDA(h, cudaStreamWaitEvent(h->chunk_stream, h->ev_chunk_fork, 0));
        int k = 0;
        for (int z0 = 0; z0 < pl.n[2]; z0 += zc, ++k) {
            const int nzc = pl.n[2] - z0 < zc ? pl.n[2] - z0 : zc;
            C* ring = D + (size_t)(k & 1) * slot;
            if (k >= 2) SCB_CUDA(h, cudaStreamWaitEvent(h->chunk_stream, h->ev_chunk_free[k & 1], 0));   // slot consumed
            LinesParams<T> p = pb2;
            p.in = Cc + (size_t)z0 * pl.PX * pl.L[1];
            p.out = ring;
            p.out_scomp = (long long)(plane * zc);
            SCB_CUDA(h, launch_lines<T>(pl.L[1], +1, p, nzc, nc, h->chunk_stream));
            SCB_CUDA(h, cudaEventRecord(h->ev_chunk_ready[k & 1], h->chunk_stream));
            SCB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_chunk_ready[k & 1], 0));
            XParams<T> q = pb3;
            q.in = ring;
            q.out = efield + (size_t)z0 * pl.n[0] * pl.n[1];
            q.nlines = (long long)pl.n[1] * nzc;
            q.cplx_scomp = (long long)(plane * zc);
            SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], q, 3, h->stream));
            if (phi) {
                q.in = ring + 3 * plane * zc;
                q.out = phi + (size_t)z0 * pl.n[0] * pl.n[1];
                SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], q, 1, h->stream));
            }
            SCB_CUDA(h, cudaEventRecord(h->ev_chunk_free[k & 1], h->stream));
        }
        tick(h, 12);   // (chunked form: B2 and B3 interleave; pass_ms[3] holds both and pass_ms[4] is ~0)
        h->launches += 2 * k - 2 + (phi ? k : 0);
    } else {
        LinesParams<T> p = pb2;
        p.in = Cc;
        p.out = D;
        p.out_scomp = (long long)szA;
        SCB_CUDA(h, launch_lines<T>(pl.L[1], +1, p, pl.n[2], nc, h->stream));
        tick(h, 12);
        XParams<T> q = pb3;
        q.in = D;
        q.out = efield;
        q.nlines = (long long)pl.n[1] * pl.n[2];
        q.cplx_scomp = (long long)szA;
        SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], q, 3, h->stream));
        if (phi) {  // fourth component -> its own output array, same FPEI / M factor
            q.in = D + 3 * szA;
            q.out = phi;
            SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], q, 1, h->stream));
            h->launches += 1;
        }
    }
    tick(h, 13);
    h->t_pass = true;
    h->t_coll = false;
    h->launches += 5;
    return SCB_OK;
}

int solve_dispatch(scb_handle* h, const void* rho, void* efield, int mdt, const int64_t n[3], const double delta[3],
                   double gamma, int mode, const double offset[3], void* phi = nullptr) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!rho || !efield || !delta) return fail(h, SCB_ERR_INVALID_ARG, "null pointer argument");
    if (!valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad mesh dtype");
    SCB_TRY(check_grid(h, n));
    for (int a = 0; a < 3; ++a)
        if (!(delta[a] > 0.0)) return fail(h, SCB_ERR_INVALID_ARG, "delta must be positive");
    if (!(gamma > 0.0)) return fail(h, SCB_ERR_INVALID_ARG, "gamma must be positive");
    SCB_CUDA(h, cudaSetDevice(h->device));
    const Plan pl = make_plan(n);
    tick(h, 2);
    int rc;
    if (mdt == SCB_F64) rc = run_solve<double>(h, (const double*)rho, (double*)efield, pl, delta, gamma, mode, offset, (double*)phi);
    else rc = run_solve<float>(h, (const float*)rho, (float*)efield, pl, delta, gamma, mode, offset, (float*)phi);
    tick(h, 3);
    h->t_solve = rc == SCB_OK;
    return rc;
}

Geom3 make_geom(const int64_t n[3], const double lo[3], const double delta[3]) {
    Geom3 g;
    static const int keep = [] { const char* e = std::getenv("SCB_L2_HINT"); return e ? std::atoi(e) : 1; }();
    g.l2_keep = keep;
    for (int a = 0; a < 3; ++a) {
        g.n[a] = (int)n[a];
        g.lo[a] = lo[a];
        g.delta[a] = delta[a];
        g.rinv[a] = 1.0 / delta[a];
        g.rinvf[a] = 1.0f / (float)delta[a];
    }
    return g;
}

double image_offset_z(int mdt, double min_z, double max_z) {
    // offset_z = 2*min_z + (max_z - min_z) in the mesh precision (src/solvers/free_space.jl:39)
    if (mdt == SCB_F32) {
        const float lo = (float)min_z, hi = (float)max_z;
        return (double)(2.0f * lo + (hi - lo));
    }
    return 2.0 * min_z + (max_z - min_z);
}

bool green_overlap_enabled() {
    static const bool on = [] { const char* e = std::getenv("SCB_GREEN_OVERLAP"); return !e || std::atoi(e) != 0; }();
    return on;
}

bool green_cached(const scb_handle* h, const GreenKey& key) {
    for (const auto& e : h->green)
        if (e.key == key && e.ncomp >= 3) return true;
    return false;
}

// Fused steps on a NEW geometry (a tracking loop re-fits the mesh to the bunch every step, so the cached spectrum
// misses every time): build the Green spectrum on a second, high-priority stream while the deposit runs on the main
// stream.  The two stress different units -- the point-wise IGF is FP64-bound and the spectrum passes stream through
// HBM, the deposit is bound by the L2 reduction path and uses a third of the DRAM bandwidth -- so most of the build
// hides behind the deposit.  The build only touches the arena and the spectrum buffers, the deposit only rho and the
// tile accumulator.  No-op when the spectrum is cached (the warm path) or the cache is disabled.
int prefetch_green(scb_handle* h, int mdt, const int64_t n[3], const double min_bounds[3], const double max_bounds[3],
                   const double delta[3], double gamma, int at_cathode) {
    if (!green_overlap_enabled() || !h->opt.green_cache || h->green_pending) return SCB_OK;
    if (!valid_dt(mdt) || !n || !min_bounds || !max_bounds || !delta || !(gamma > 0.0)) return SCB_OK;   // the solve reports it
    for (int a = 0; a < 3; ++a)
        if (n[a] < 2 || n[a] > kMaxFftLen / 2 || !(delta[a] > 0.0)) return SCB_OK;
    const Plan pl = make_plan(n);
    const double zero3[3] = {0, 0, 0};
    double offset[3] = {0.0, 0.0, 0.0};
    if (at_cathode) offset[2] = image_offset_z(mdt, min_bounds[2], max_bounds[2]);
    const GreenKey kfree = make_key(pl, delta, gamma, zero3, mdt, 0);
    const GreenKey kimg = make_key(pl, delta, gamma, offset, mdt, 1);
    if (green_cached(h, kfree) && (!at_cathode || green_cached(h, kimg))) return SCB_OK;
    if (!h->green_stream) {
        int lo = 0, hi = 0;
        SCB_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SCB_CUDA(h, cudaStreamCreateWithPriority(&h->green_stream, cudaStreamNonBlocking, hi));
        SCB_CUDA(h, cudaEventCreateWithFlags(&h->ev_green_start, cudaEventDisableTiming));
        SCB_CUDA(h, cudaEventCreateWithFlags(&h->ev_green_done, cudaEventDisableTiming));
    }
    // everything queued so far (the previous solve's passes use the arena) comes first
    SCB_CUDA(h, cudaEventRecord(h->ev_green_start, h->stream));
    SCB_CUDA(h, cudaStreamWaitEvent(h->green_stream, h->ev_green_start, 0));
    cudaStream_t main_stream = h->stream;
    const bool timing = h->timing;
    h->stream = h->green_stream;   // the build launches on h->stream
    h->timing = false;             // its events would land on the side stream
    const GreenEntry* unused = nullptr;
    int rc = get_green(h, pl, kfree, &unused);
    if (rc == SCB_OK && at_cathode) rc = get_green(h, pl, kimg, &unused);
    h->stream = main_stream;
    h->timing = timing;
    // Recorded even when a build failed half-way: whatever WAS built and cached (e.g. the free-space spectrum when the
    // image spectrum failed) was built on green_stream, and the next solve that finds it in the cache has to wait for it.
    const cudaError_t ev = cudaEventRecord(h->ev_green_done, h->green_stream);
    if (ev == cudaSuccess) h->green_pending = true;
    if (rc != SCB_OK) return rc;
    if (ev != cudaSuccess) return cuda_fail(h, ev, "cudaEventRecord(ev_green_done)");
    return SCB_OK;
}

double key_to_double(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}

}  // namespace

// =============================================================================================
extern "C" {

int scb_version(void) { return 100; }

int scb_create(int device, void* cuda_stream, const scb_options* opt, scb_handle** out) {
    if (!out) return SCB_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return SCB_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SCB_ERR_NO_DEVICE;
    if (prop.major != 10) return SCB_ERR_NO_DEVICE;  // sm_100a code only; no fallback path exists
    if (cudaSetDevice(device) != cudaSuccess) return SCB_ERR_CUDA;
    scb_handle* h = new scb_handle();
    h->device = device;
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    h->opt.green_cache = 1;
    if (opt) h->opt = *opt;
    if (const char* e = std::getenv("SCB_DEPOSIT_MODE")) h->opt.deposit_mode = std::atoi(e);  // tuning knob
    if (const char* e = std::getenv("SCB_PARTICLE_ORDER")) h->opt.particle_order = std::atoi(e);
    if (h->opt.particle_order < SCB_ORDER_RANDOM || h->opt.particle_order > SCB_ORDER_AUTO) h->opt.particle_order = SCB_ORDER_RANDOM;
    if (cudaMalloc(&h->d_bounds, 6 * sizeof(unsigned long long)) != cudaSuccess) {
        delete h;
        return SCB_ERR_ALLOC;
    }
    *out = h;
    return SCB_OK;
}

int scb_destroy(scb_handle* h) {
    if (!h) return SCB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (cudaEvent_t e : h->ev_bcast) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_pack) if (e) cudaEventDestroy(e);
    if (h->ev_field) cudaEventDestroy(h->ev_field);
    for (cudaEvent_t e : {h->ev_green_start, h->ev_green_done}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_chunk_fork, h->ev_chunk_ready[0], h->ev_chunk_ready[1], h->ev_chunk_free[0], h->ev_chunk_free[1]})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t cs : {h->copy_stream, h->d2h_stream, h->comm_stream, h->pack_stream, h->green_stream, h->chunk_stream})
        if (cs) {
            cudaStreamSynchronize(cs);
            cudaStreamDestroy(cs);
        }
    free_green(h, true);
    for (auto& kv : h->twiddles) cudaFree(kv.second);
    if (h->arena) cudaFree(h->arena);
    if (h->sh_arena) cudaFree(h->sh_arena);
    for (void* st : h->stage)
        if (st) cudaFree(st);
    if (h->packed) cudaFree(h->packed);
    if (h->tiles) cudaFree(h->tiles);
    if (h->slab) cudaFree(h->slab);
    close_peers(h);
    if (h->d_ipc) cudaFree(h->d_ipc);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->d_bounds) cudaFree(h->d_bounds);
    if (h->ev_ready)
        for (auto& e : h->ev) cudaEventDestroy(e);
    for (auto& v : h->chunk_ev)
        for (auto& e : v) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->slot_in_free[0], h->slot_in_free[1], h->slot_out_free[0], h->slot_out_free[1]})
        if (e) cudaEventDestroy(e);
    delete h;
    return SCB_OK;
}

int scb_set_stream(scb_handle* h, void* cuda_stream) {
    if (!h) return SCB_ERR_INVALID_ARG;
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return SCB_OK;
}

int scb_sync(scb_handle* h) {
    if (!h) return SCB_ERR_INVALID_ARG;
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    return SCB_OK;
}

const char* scb_last_error(const scb_handle* h) { return h ? h->err.c_str() : "invalid handle"; }

int scb_enable_timing(scb_handle* h, int on) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (on && !h->ev_ready) {
        for (auto& e : h->ev) SCB_CUDA(h, cudaEventCreate(&e));
        h->ev_ready = true;
    }
    h->timing = on != 0;
    return SCB_OK;
}

int scb_get_timing(scb_handle* h, scb_timing* out) {
    if (!h || !out) return SCB_ERR_INVALID_ARG;
    if (!h->ev_ready) return fail(h, SCB_ERR_INVALID_ARG, "timing was never enabled");
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    auto el = [&](int a, int b) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {
            (void)cudaGetLastError();
            return 0.f;
        }
        return ms;
    };
    if (h->t_dep) h->last.deposit_ms = el(0, 1);
    if (h->t_solve) h->last.solve_ms = el(2, 3);
    if (h->t_interp) h->last.interpolate_ms = el(4, 5);
    h->last.green_ms = h->t_green ? el(6, 7) : 0.f;
    if (h->t_pass)
        for (int i = 0; i < 5; ++i) h->last.pass_ms[i] = el(8 + i, 9 + i);
    h->last.pass_ms[5] = h->last.pass_ms[6] = 0.f;
    if (h->t_pass && h->t_coll) {   // slab-decomposed solve: the collectives that pass_ms[0] and pass_ms[4] include
        h->last.pass_ms[5] = el(8, 14);    // reduce-scatter of rho
        h->last.pass_ms[6] = el(15, 13);   // all-gather of E
    }
    *out = h->last;
    return SCB_OK;
}

int64_t scb_launch_count(const scb_handle* h) { return h ? h->launches : 0; }

int scb_clear(scb_handle* h, void* rho, const int64_t n[3], int mdt) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!rho || !n || !valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_clear");
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_CUDA(h, cudaMemsetAsync(rho, 0, (size_t)n[0] * n[1] * n[2] * dt_size(mdt), h->stream));
    return SCB_OK;
}

int scb_deposit(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q, int pdt,
                void* rho, int mdt, const int64_t n[3], const double min_bounds[3], const double delta[3], int clear) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !rho || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_deposit");
    if (np > 0 && (!x || !y || !z || !q)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    tick(h, 0);
    SCB_TRY(run_deposit(h, np, x, y, z, q, pdt, rho, mdt, make_geom(n, min_bounds, delta), clear != 0, false));
    tick(h, 1);
    h->t_dep = true;
    return SCB_OK;
}

int scb_solve(scb_handle* h, const void* rho, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
              const double max_bounds[3], const double delta[3], double gamma, int at_cathode) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!min_bounds || !max_bounds) return fail(h, SCB_ERR_INVALID_ARG, "null bounds");
    double offset[3] = {0.0, 0.0, 0.0};
    if (at_cathode) {
        if (!valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad mesh dtype");
        offset[2] = image_offset_z(mdt, min_bounds[2], max_bounds[2]);
    }
    return solve_dispatch(h, rho, efield, mdt, n, delta, gamma, at_cathode ? 1 : 0, offset);
}

int scb_solve_freespace(scb_handle* h, const void* rho, void* efield, int mdt, const int64_t n[3],
                        const double delta[3], double gamma, const double offset[3]) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!offset) return fail(h, SCB_ERR_INVALID_ARG, "null offset");
    const bool zero = offset[0] == 0.0 && offset[1] == 0.0 && offset[2] == 0.0;
    return solve_dispatch(h, rho, efield, mdt, n, delta, gamma, zero ? 0 : 2, offset);
}

int scb_solve_potential(scb_handle* h, const void* rho, void* efield, void* phi, int mdt, const int64_t n[3],
                        const double min_bounds[3], const double max_bounds[3], const double delta[3], double gamma,
                        int at_cathode) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!min_bounds || !max_bounds || !phi) return fail(h, SCB_ERR_INVALID_ARG, "null bounds or phi");
    double offset[3] = {0.0, 0.0, 0.0};
    if (at_cathode) {
        if (!valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad mesh dtype");
        offset[2] = image_offset_z(mdt, min_bounds[2], max_bounds[2]);
    }
    return solve_dispatch(h, rho, efield, mdt, n, delta, gamma, at_cathode ? 1 : 0, offset, phi);
}

int scb_bfield(scb_handle* h, const void* efield, void* bfield, int mdt, const int64_t n[3], double gamma) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!efield || !bfield || !valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_bfield");
    if (!(gamma >= 1.0)) return fail(h, SCB_ERR_INVALID_ARG, "gamma must be >= 1");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    // beta / c for a bunch moving along +z; formed in double, applied in the mesh precision
    const double beta_over_c = std::sqrt(1.0 - 1.0 / (gamma * gamma)) / kCLight;
    SCB_CUDA(h, launch_bfield(mdt, efield, bfield, (long long)n[0] * n[1] * n[2], beta_over_c, h->stream));
    h->launches += 1;
    return SCB_OK;
}

int scb_interpolate(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt,
                    const void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                    const double delta[3], void* ex, void* ey, void* ez) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !efield || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_interpolate");
    if (np > 0 && (!x || !y || !z || !ex || !ey || !ez)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    tick(h, 4);
    if (np > 0) SCB_TRY(run_interpolate(h, np, x, y, z, pdt, efield, mdt, make_geom(n, min_bounds, delta), ex, ey, ez, nullptr));
    tick(h, 5);
    h->t_interp = true;
    return SCB_OK;
}

int scb_interpolate_kick(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt,
                         const void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                         const double delta[3], void* px, void* py, void* pz, double coef_xy, double coef_z) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !efield || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_interpolate_kick");
    if (np > 0 && (!x || !y || !z || !px || !py || !pz)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    Kick k;
    k.on = 1;
    k.cxy = coef_xy;
    k.cz = coef_z;
    tick(h, 4);
    if (np > 0) SCB_TRY(run_interpolate(h, np, x, y, z, pdt, efield, mdt, make_geom(n, min_bounds, delta), px, py, pz, nullptr, k));
    tick(h, 5);
    h->t_interp = true;
    return SCB_OK;
}

int scb_green(scb_handle* h, void* out, const int64_t n2[3], const double delta[3], double gamma, int icomp,
              const double offset[3], int dt) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!out || !n2 || !delta || !offset || !valid_dt(dt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_green");
    for (int a = 0; a < 3; ++a)
        if (n2[a] < 2 || n2[a] > 2 * kMaxFftLen) return fail(h, SCB_ERR_INVALID_ARG, "bad doubled grid size");
    SCB_CUDA(h, cudaSetDevice(h->device));
    IgfGeom g{};
    for (int a = 0; a < 3; ++a) {
        g.isize[a] = (int)n2[a];
        g.i0[a] = 1;
        g.cnt[a] = (int)n2[a];
        g.delta[a] = delta[a];
        g.offset[a] = offset[a];
    }
    g.gamma = gamma;
    const size_t total = (size_t)n2[0] * n2[1] * n2[2];
    SCB_TRY(ensure_arena(h, total * 8));
    double* P = static_cast<double*>(h->arena);
    SCB_CUDA(h, launch_green_point(P, g, icomp, h->stream));
    SCB_CUDA(h, launch_green_reference_layout(out, dt == SCB_F64, P, (int)n2[0], (int)n2[1], (int)n2[2], h->stream));
    h->launches += 2;
    return SCB_OK;
}

int scb_bounds(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt, double out_min[3],
               double out_max[3]) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np <= 0) return fail(h, SCB_ERR_INVALID_ARG, "Particle arrays cannot be empty.");
    if (!x || !y || !z || !out_min || !out_max || !valid_dt(pdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_bounds");
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_CUDA(h, launch_bounds(pdt, np, x, y, z, reinterpret_cast<double*>(h->d_bounds), h->stream));
    h->launches += 2;
    unsigned long long host[6];
    SCB_CUDA(h, cudaMemcpyAsync(host, h->d_bounds, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int a = 0; a < 3; ++a) {
        out_min[a] = key_to_double(host[a]);
        out_max[a] = key_to_double(host[3 + a]);
    }
    return SCB_OK;
}

int scb_cell_index(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt, int mdt,
                   const double min_bounds[3], const double delta[3], int64_t* ix, int64_t* iy, int64_t* iz) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_cell_index");
    if (np > 0 && (!x || !y || !z || !ix || !iy || !iz)) return fail(h, SCB_ERR_INVALID_ARG, "null array");
    SCB_CUDA(h, cudaSetDevice(h->device));
    const int64_t n1[3] = {2, 2, 2};
    SCB_CUDA(h, launch_cell_index(pdt, mdt, np, x, y, z, make_geom(n1, min_bounds, delta), (long long*)ix,
                                  (long long*)iy, (long long*)iz, h->stream));
    if (np > 0) h->launches += 1;
    return SCB_OK;
}

// ---- bunches kept ordered by cell -----------------------------------------------------------------
int scb_set_particle_order(scb_handle* h, int order) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (order < SCB_ORDER_RANDOM || order > SCB_ORDER_AUTO) return fail(h, SCB_ERR_INVALID_ARG, "unknown particle order");
    h->opt.particle_order = order;
    h->auto_hold = 0;   // SCB_ORDER_AUTO probes on its next deposit
    return SCB_OK;
}

int scb_sort_particles(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt, int mdt,
                       const int64_t n[3], const double min_bounds[3], const double delta[3], uint32_t* perm_out) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || np >= (int64_t)1 << 31 || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_sort_particles");
    if (np > 0 && (!x || !y || !z || !perm_out)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    if (np == 0) return SCB_OK;
    SCB_CUDA(h, cudaSetDevice(h->device));
    const long long ncell = (long long)n[0] * n[1] * n[2];
    int key_bits = 1;
    while (key_bits < 32 && (1ll << key_bits) < ncell) ++key_bits;
    SCB_TRY(ensure_arena(h, sort_scratch_bytes(np)));
    SCB_CUDA(h, launch_cell_keys(pdt, mdt, np, x, y, z, make_geom(n, min_bounds, delta), h->arena, key_bits, h->stream));
    int launches = 0;
    SCB_CUDA(h, launch_sort_pairs(h->arena, np, key_bits, perm_out, h->stream, &launches));
    h->launches += 1 + launches;
    return SCB_OK;
}

int scb_permute(scb_handle* h, int64_t np, const uint32_t* perm, int nfields, const void* const* src, void* const* dst, int dt) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || nfields < 0 || nfields > 8 || !valid_dt(dt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_permute");
    if (np == 0 || nfields == 0) return SCB_OK;
    if (!perm || !src || !dst) return fail(h, SCB_ERR_INVALID_ARG, "null array");
    for (int f = 0; f < nfields; ++f)
        if (!src[f] || !dst[f] || src[f] == dst[f]) return fail(h, SCB_ERR_INVALID_ARG, "scb_permute: null or aliased field");
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_CUDA(h, launch_permute((int)dt_size(dt), np, perm, nfields, src, dst, h->stream));
    h->launches += 1;
    return SCB_OK;
}

int scb_particle_order_fraction(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, int pdt, int mdt,
                                const int64_t n[3], const double min_bounds[3], const double delta[3], double* fraction_out) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !min_bounds || !delta || !fraction_out || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_particle_order_fraction");
    if (np > 0 && (!x || !y || !z)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_CUDA(h, launch_order_probe(pdt, mdt, np, x, y, z, make_geom(n, min_bounds, delta), h->d_bounds, h->stream));
    h->launches += 1;
    unsigned long long host[2];
    SCB_CUDA(h, cudaMemcpyAsync(host, h->d_bounds, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    *fraction_out = host[1] ? (double)host[0] / (double)host[1] : 0.0;
    return SCB_OK;
}

int scb_step(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q, int pdt,
             void* rho, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
             const double max_bounds[3], const double delta[3], double gamma, int at_cathode, void* ex, void* ey,
             void* ez) {
    if (h && np > 0) SCB_TRY(prefetch_green(h, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    SCB_TRY(scb_deposit(h, np, x, y, z, q, pdt, rho, mdt, n, min_bounds, delta, 1));
    SCB_TRY(scb_solve(h, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    return scb_interpolate(h, np, x, y, z, pdt, efield, mdt, n, min_bounds, delta, ex, ey, ez);
}

// ---- strided / AoS particle layouts ------------------------------------------
namespace {
// element strides from the C struct; `need_q` / `need_out`: which members the call reads
int layout_from(scb_handle* h, const scb_particle_strides* st, bool need_q, bool need_out, PLayout* L) {
    if (!st) return fail(h, SCB_ERR_INVALID_ARG, "null strides");
    if (st->x < 1 || st->y < 1 || st->z < 1) return fail(h, SCB_ERR_INVALID_ARG, "coordinate strides must be >= 1");
    if (need_q && st->q < 0) return fail(h, SCB_ERR_INVALID_ARG, "charge stride must be >= 0 (0 = one charge for all particles)");
    if (need_out && (st->ex < 1 || st->ey < 1 || st->ez < 1)) return fail(h, SCB_ERR_INVALID_ARG, "output strides must be >= 1");
    L->x = st->x; L->y = st->y; L->z = st->z;
    L->q = need_q ? st->q : 1;
    if (need_out) { L->ex = st->ex; L->ey = st->ey; L->ez = st->ez; }
    return SCB_OK;
}
}  // namespace

int scb_deposit_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q,
                        const scb_particle_strides* st, int pdt, void* rho, int mdt, const int64_t n[3],
                        const double min_bounds[3], const double delta[3], int clear) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !rho || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_deposit_strided");
    if (np > 0 && (!x || !y || !z || !q)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    PLayout L;
    SCB_TRY(layout_from(h, st, true, false, &L));
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    tick(h, 0);
    SCB_TRY(run_deposit(h, np, x, y, z, q, pdt, rho, mdt, make_geom(n, min_bounds, delta), clear != 0, false, &L));
    tick(h, 1);
    h->t_dep = true;
    return SCB_OK;
}

int scb_interpolate_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                            const scb_particle_strides* st, int pdt, const void* efield, int mdt, const int64_t n[3],
                            const double min_bounds[3], const double delta[3], void* ex, void* ey, void* ez) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !efield || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_interpolate_strided");
    if (np > 0 && (!x || !y || !z || !ex || !ey || !ez)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    PLayout L;
    SCB_TRY(layout_from(h, st, false, true, &L));
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    tick(h, 4);
    if (np > 0) SCB_TRY(run_interpolate(h, np, x, y, z, pdt, efield, mdt, make_geom(n, min_bounds, delta), ex, ey, ez, nullptr, Kick(), &L));
    tick(h, 5);
    h->t_interp = true;
    return SCB_OK;
}

int scb_interpolate_kick_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                                 const scb_particle_strides* st, int pdt, const void* efield, int mdt,
                                 const int64_t n[3], const double min_bounds[3], const double delta[3], void* px,
                                 void* py, void* pz, double coef_xy, double coef_z) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np < 0 || !efield || !min_bounds || !delta || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_interpolate_kick_strided");
    if (np > 0 && (!x || !y || !z || !px || !py || !pz)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    PLayout L;
    SCB_TRY(layout_from(h, st, false, true, &L));
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    Kick k;
    k.on = 1;
    k.cxy = coef_xy;
    k.cz = coef_z;
    tick(h, 4);
    if (np > 0) SCB_TRY(run_interpolate(h, np, x, y, z, pdt, efield, mdt, make_geom(n, min_bounds, delta), px, py, pz, nullptr, k, &L));
    tick(h, 5);
    h->t_interp = true;
    return SCB_OK;
}

int scb_bounds_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z,
                       const scb_particle_strides* st, int pdt, double out_min[3], double out_max[3]) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (np <= 0) return fail(h, SCB_ERR_INVALID_ARG, "Particle arrays cannot be empty.");
    if (!x || !y || !z || !out_min || !out_max || !valid_dt(pdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_bounds_strided");
    PLayout L;
    SCB_TRY(layout_from(h, st, false, false, &L));
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_CUDA(h, launch_bounds(pdt, np, x, y, z, reinterpret_cast<double*>(h->d_bounds), h->stream, &L));
    h->launches += 2;
    unsigned long long host[6];
    SCB_CUDA(h, cudaMemcpyAsync(host, h->d_bounds, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int a = 0; a < 3; ++a) {
        out_min[a] = key_to_double(host[a]);
        out_max[a] = key_to_double(host[3 + a]);
    }
    return SCB_OK;
}

int scb_step_strided(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q,
                     const scb_particle_strides* st, int pdt, void* rho, void* efield, int mdt, const int64_t n[3],
                     const double min_bounds[3], const double max_bounds[3], const double delta[3], double gamma,
                     int at_cathode, void* ex, void* ey, void* ez) {
    if (h && np > 0) SCB_TRY(prefetch_green(h, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    SCB_TRY(scb_deposit_strided(h, np, x, y, z, q, st, pdt, rho, mdt, n, min_bounds, delta, 1));
    SCB_TRY(scb_solve(h, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    return scb_interpolate_strided(h, np, x, y, z, st, pdt, efield, mdt, n, min_bounds, delta, ex, ey, ez);
}

// comm_mode 0: one GPU; 1: particle shards, rho all-reduced, solve replicated; 2: particle shards, slab-decomposed solve
static int step_host_impl(scb_handle* h, int64_t np, const void* xh, const void* yh, const void* zh, const void* qh, int pdt,
                          void* rho, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                          const double max_bounds[3], const double delta[3], double gamma, int at_cathode, void* exh,
                          void* eyh, void* ezh, int comm_mode) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (comm_mode != 0 && !h->comm) return fail(h, SCB_ERR_COMM, "scb_comm_init has not been called");
    if (np <= 0 || !xh || !yh || !zh || !qh || !exh || !eyh || !ezh || !rho || !efield || !valid_dt(pdt) || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_step_host");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    const size_t es = dt_size(pdt);
    const size_t arr = ((size_t)np * es + 255) / 256 * 256;
    const int slot = h->host_steps_in_flight & 1;
    if (h->stage_bytes[slot] < 7 * arr) {
        if (h->stage[slot]) {
            SCB_CUDA(h, cudaDeviceSynchronize());
            cudaFree(h->stage[slot]);
            h->stage[slot] = nullptr;
            h->stage_bytes[slot] = 0;
            h->slot_used[slot] = false;
        }
        if (cudaMalloc(&h->stage[slot], 7 * arr) != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(h, SCB_ERR_ALLOC, "particle staging allocation failed");
        }
        h->stage_bytes[slot] = 7 * arr;
    }
    if (!h->copy_stream) SCB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->d2h_stream) SCB_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&h->slot_in_free[slot], &h->slot_out_free[slot]})
        if (!*e) SCB_CUDA(h, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    char* d[7];
    for (int i = 0; i < 7; ++i) d[i] = static_cast<char*>(h->stage[slot]) + i * arr;
    const void* src[4] = {xh, yh, zh, qh};
    void* dsth[3] = {exh, eyh, ezh};

    const int64_t chunk = 1 << 22;
    const int nchunk = (int)((np + chunk - 1) / chunk);
    std::vector<cudaEvent_t>& cev = h->chunk_ev[slot];
    while ((int)cev.size() < 2 * nchunk + 1) {
        cudaEvent_t e;
        SCB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cev.push_back(e);
    }
    const Geom3 g = make_geom(n, min_bounds, delta);
    // hides behind the upload (the slab solve keeps its own slab of the spectrum and builds it inside the solve)
    if (comm_mode != 2) SCB_TRY(prefetch_green(h, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    if (h->host_steps_in_flight == 0) {
        // first step of a pipeline: the upload waits for whatever the caller queued before this call
        SCB_CUDA(h, cudaEventRecord(cev[2 * nchunk], h->stream));
        SCB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, cev[2 * nchunk], 0));
    }
    // the slot's inputs were last read by the gather of the step two calls back
    if (h->slot_used[slot]) SCB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->slot_in_free[slot], 0));
    SCB_CUDA(h, cudaMemsetAsync(rho, 0, (size_t)n[0] * n[1] * n[2] * dt_size(mdt), h->stream));
    for (int c = 0; c < nchunk; ++c) {
        const int64_t o = (int64_t)c * chunk, m = (np - o) < chunk ? (np - o) : chunk;
        for (int a = 0; a < 4; ++a)
            SCB_CUDA(h, cudaMemcpyAsync(d[a] + o * es, (const char*)src[a] + o * es, m * es, cudaMemcpyHostToDevice, h->copy_stream));
        SCB_CUDA(h, cudaEventRecord(cev[c], h->copy_stream));
        SCB_CUDA(h, cudaStreamWaitEvent(h->stream, cev[c], 0));
        SCB_CUDA(h, launch_deposit(pdt, mdt, m, d[0] + o * es, d[1] + o * es, d[2] + o * es, d[3] + o * es, rho, g, h->opt.deposit_mode, h->stream));
        h->launches += 1;
    }
    if (comm_mode == 2) {
        SCB_TRY(scb_solve_sharded(h, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    } else {
        if (comm_mode == 1)
            SCB_NCCL(h, g_nccl.AllReduce(rho, rho, (size_t)n[0] * n[1] * n[2], mdt == SCB_F64 ? ncclFloat64 : ncclFloat32,
                                         ncclSum, h->comm, h->stream));
        SCB_TRY(scb_solve(h, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
    }
    bool packed_ready = false;
    if (np * 4 >= (int64_t)n[0] * n[1] * n[2]) {
        SCB_TRY(ensure_packed(h, (size_t)n[0] * n[1] * n[2] * packed_bytes_per_node(mdt)));
        SCB_CUDA(h, launch_pack_efield(mdt, efield, h->packed, g, h->stream));
        h->launches += 1;
        packed_ready = true;
    }
    // the slot's outputs were last read by the download of the step two calls back
    if (h->slot_used[slot]) SCB_CUDA(h, cudaStreamWaitEvent(h->stream, h->slot_out_free[slot], 0));
    for (int c = 0; c < nchunk; ++c) {
        const int64_t o = (int64_t)c * chunk, m = (np - o) < chunk ? (np - o) : chunk;
        if (packed_ready) {
            SCB_CUDA(h, launch_interpolate_packed(pdt, mdt, m, d[0] + o * es, d[1] + o * es, d[2] + o * es, h->packed, g,
                                                  d[4] + o * es, d[5] + o * es, d[6] + o * es, h->stream));
        } else {
            SCB_CUDA(h, launch_interpolate(pdt, mdt, m, d[0] + o * es, d[1] + o * es, d[2] + o * es, efield, g, d[4] + o * es,
                                           d[5] + o * es, d[6] + o * es, h->stream));
        }
        h->launches += 1;
        SCB_CUDA(h, cudaEventRecord(cev[nchunk + c], h->stream));
        SCB_CUDA(h, cudaStreamWaitEvent(h->d2h_stream, cev[nchunk + c], 0));
        for (int a = 0; a < 3; ++a)
            SCB_CUDA(h, cudaMemcpyAsync((char*)dsth[a] + o * es, d[4 + a] + o * es, m * es, cudaMemcpyDeviceToHost, h->d2h_stream));
    }
    SCB_CUDA(h, cudaEventRecord(h->slot_in_free[slot], h->stream));
    SCB_CUDA(h, cudaEventRecord(h->slot_out_free[slot], h->d2h_stream));
    h->slot_used[slot] = true;
    h->host_steps_in_flight += 1;
    return SCB_OK;
}

int scb_step_host_async(scb_handle* h, int64_t np, const void* xh, const void* yh, const void* zh, const void* qh, int pdt,
                        void* rho, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                        const double max_bounds[3], const double delta[3], double gamma, int at_cathode, void* exh,
                        void* eyh, void* ezh) {
    return step_host_impl(h, np, xh, yh, zh, qh, pdt, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode,
                          exh, eyh, ezh, 0);
}

int scb_step_host_sharded_async(scb_handle* h, int64_t np, const void* xh, const void* yh, const void* zh, const void* qh,
                                int pdt, void* rho_partial, void* efield, int mdt, const int64_t n[3],
                                const double min_bounds[3], const double max_bounds[3], const double delta[3], double gamma,
                                int at_cathode, int slab_solve, void* exh, void* eyh, void* ezh) {
    return step_host_impl(h, np, xh, yh, zh, qh, pdt, rho_partial, efield, mdt, n, min_bounds, max_bounds, delta, gamma,
                          at_cathode, exh, eyh, ezh, slab_solve ? 2 : 1);
}

int scb_step_host_wait(scb_handle* h) {
    if (!h) return SCB_ERR_INVALID_ARG;
    SCB_CUDA(h, cudaSetDevice(h->device));
    if (h->copy_stream) SCB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    if (h->d2h_stream) SCB_CUDA(h, cudaStreamSynchronize(h->d2h_stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->host_steps_in_flight = 0;
    return SCB_OK;
}

int scb_step_host(scb_handle* h, int64_t np, const void* xh, const void* yh, const void* zh, const void* qh, int pdt,
                  void* rho, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                  const double max_bounds[3], const double delta[3], double gamma, int at_cathode, void* exh,
                  void* eyh, void* ezh) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (h->host_steps_in_flight) SCB_TRY(scb_step_host_wait(h));
    const int rc = scb_step_host_async(h, np, xh, yh, zh, qh, pdt, rho, efield, mdt, n, min_bounds, max_bounds, delta, gamma,
                                       at_cathode, exh, eyh, ezh);
    const int rw = scb_step_host_wait(h);
    return rc != SCB_OK ? rc : rw;
}

int scb_drop_green_cache(scb_handle* h) {
    if (!h) return SCB_ERR_INVALID_ARG;
    free_green(h, false);
    return SCB_OK;
}

int64_t scb_workspace_bytes(const scb_handle* h) {
    if (!h) return 0;
    int64_t b = (int64_t)h->arena_bytes + (int64_t)h->stage_bytes[0] + (int64_t)h->stage_bytes[1] + (int64_t)h->packed_bytes + (int64_t)h->tiles_bytes;
    for (auto& e : h->green) b += (int64_t)e.cap;
    for (auto& e : h->green_pool) b += (int64_t)e.second;
    return b;
}

}  // extern "C"

// =============================================================================================
// multi-GPU: NCCL is loaded lazily so that the library has no link-time dependency on it
namespace {

// Cross-rank barrier in stream order: a one-int all-reduce completes on a rank only after every rank
// has reached it, i.e. after every rank's preceding kernels (and their peer stores) have finished.
int rank_barrier(scb_handle* h) {
    int* bar = reinterpret_cast<int*>(h->d_ipc + (size_t)SCB_MAX_RANKS * 64 + 4);
    SCB_NCCL(h, g_nccl.AllReduce(bar, bar, 1, ncclInt32, ncclSum, h->comm, h->stream));
    return SCB_OK;
}

void close_peers(scb_handle* h) {
    for (int r = 0; r < (int)h->peer_arena.size(); ++r)
        if (r != h->rank && h->peer_arena[r]) cudaIpcCloseMemHandle(h->peer_arena[r]);
    h->peer_arena.clear();
}

// (Re)map the other ranks' arenas after this rank's arena was (re)allocated.  Collective: every rank
// grows its arena in the same call, so every rank gets here together.  Any failure on any rank
// switches all ranks back to the NCCL send/recv path.
int exchange_arenas(scb_handle* h) {
    // Entered or skipped by ALL ranks together: sh_gen only changes inside run_solve_sharded (same arguments on every
    // rank), and a rank that does not want the peer path (SCB_P2P=0) still takes part and votes against it below.
    if (h->peer_gen == h->sh_gen) return SCB_OK;
    const int G = h->nranks, me = h->rank;
    int want = 1;
    if (const char* e = std::getenv("SCB_P2P")) want = std::atoi(e) != 0;
    if (!h->d_ipc) {
        SCB_CUDA(h, cudaMalloc(&h->d_ipc, (size_t)SCB_MAX_RANKS * 64 + 8));
        SCB_CUDA(h, cudaMemsetAsync(h->d_ipc, 0, (size_t)SCB_MAX_RANKS * 64 + 8, h->stream));
    }
    close_peers(h);
    int ok = want;
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, h->sh_arena) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; std::memset(&mine, 0, sizeof(mine)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SCB_CUDA(h, cudaMemcpyAsync(h->d_ipc + (size_t)me * 64, &mine, 64, cudaMemcpyHostToDevice, h->stream));
    SCB_NCCL(h, g_nccl.AllGather(h->d_ipc + (size_t)me * 64, h->d_ipc, 64, ncclChar, h->comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all(G);
    SCB_CUDA(h, cudaMemcpyAsync(all.data(), h->d_ipc, (size_t)G * 64, cudaMemcpyDeviceToHost, h->stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->peer_arena.assign(G, nullptr);
    h->peer_arena[me] = h->sh_arena;
    for (int r = 0; r < G && ok; ++r) {
        if (r == me) continue;
        if (cudaIpcOpenMemHandle(&h->peer_arena[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            (void)cudaGetLastError();
            h->peer_arena[r] = nullptr;
            ok = 0;
        }
    }
    // consensus: all ranks use the peer path or none does
    int* flag = reinterpret_cast<int*>(h->d_ipc + (size_t)SCB_MAX_RANKS * 64);
    SCB_CUDA(h, cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    SCB_NCCL(h, g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, h->comm, h->stream));
    int all_ok = 0;
    SCB_CUDA(h, cudaMemcpyAsync(&all_ok, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    SCB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->peer_gen = h->sh_gen;
    if (!all_ok) {
        close_peers(h);
        h->p2p = 0;
        return SCB_OK;
    }
    h->p2p = 1;
    return SCB_OK;
}

// all-to-all of equal contiguous blocks (block b of `send` goes to rank b; block g of `recv` comes
// from rank g), `nsets` independent buffers (field components) in one NCCL group
int all_to_all(scb_handle* h, const char* send, char* recv, size_t block_bytes, int nsets, size_t set_stride_bytes) {
    SCB_NCCL(h, g_nccl.GroupStart());
    for (int s = 0; s < nsets; ++s)
        for (int r = 0; r < h->nranks; ++r) {
            SCB_NCCL(h, g_nccl.Send(send + s * set_stride_bytes + (size_t)r * block_bytes, block_bytes, ncclChar, r, h->comm, h->stream));
            SCB_NCCL(h, g_nccl.Recv(recv + s * set_stride_bytes + (size_t)r * block_bytes, block_bytes, ncclChar, r, h->comm, h->stream));
        }
    SCB_NCCL(h, g_nccl.GroupEnd());
    return SCB_OK;
}

// Which axis the slab-decomposed solve cuts between the x pass and the y / z passes.
//   kx (default, round 2): after F1 the z slabs are exchanged for kx slabs (every rank: all y, all z, PX / ranks kx), F2,
//       the z pass and B2 run entirely locally -- the z pass with the single-GPU kernels (k_z_eo / k_z_tma) -- and the
//       y-PRUNED result goes back to z slabs for B3.  Bytes on the links per rank: (A + 3A) / G * (G-1)/G.
//   ky (round 1): F1 and F2 on z slabs, exchange for ky slabs, z pass, back.  (B + 3B) / G * (G-1)/G with B = 2A, the
//       z pass stores to peers element by element.  Needs the padded y length divisible by the ranks.
bool shard_by_kx(const scb_handle* h, const Plan& pl) {
    static const int want_ky = [] { const char* e = std::getenv("SCB_SHARD"); return e && (e[0] == 'k' && e[1] == 'y'); }();
    const bool kx_ok = pl.PX % h->nranks == 0, ky_ok = pl.L[1] % h->nranks == 0;
    return kx_ok && !(want_ky && ky_ok);
}

template <typename T>
int run_solve_sharded_kx(scb_handle* h, const T* rho_partial, T* efield, const Plan& pl, const double delta[3], double gamma,
                         int mode, const double offset[3], bool defer_gather) {
    using C = cx_t<T>;
    const int G = h->nranks, me = h->rank;
    const int mdt = sizeof(T) == 8 ? SCB_F64 : SCB_F32;
    const ncclDataType_t nt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    const int nzl = pl.n[2] / G, PXl = pl.PX / G;
    const int kx0 = me * PXl;
    const int ninl = pl.ninner - kx0 < 0 ? 0 : (pl.ninner - kx0 < PXl ? pl.ninner - kx0 : PXl);   // valid kx of this rank
    const double zero3[3] = {0, 0, 0};
    const GreenEntry* gfree = nullptr;
    const GreenEntry* gaux = nullptr;
    h->t_green = false;
    SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, zero3, mdt, 0), &gfree));
    if (mode == 1) {
        SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, offset, mdt, 1), &gaux));
        for (auto& e : h->green)
            if (e.key == make_key(pl, delta, gamma, zero3, mdt, 0)) gfree = &e;
    }
    if (h->green_pending) {
        SCB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_green_done, 0));
        h->green_pending = false;
    }
    const size_t slab_elems = (size_t)pl.n[0] * pl.n[1] * nzl;
    if (h->slab_bytes < slab_elems * sizeof(T)) {
        if (h->slab) { SCB_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(h->slab); h->slab = nullptr; h->slab_bytes = 0; }
        if (cudaMalloc(&h->slab, slab_elems * sizeof(T)) != cudaSuccess) { (void)cudaGetLastError(); return fail(h, SCB_ERR_ALLOC, "slab allocation failed"); }
        h->slab_bytes = slab_elems * sizeof(T);
    }
    T* slab = static_cast<T*>(h->slab);
    const size_t a1 = (size_t)PXl * pl.n[1] * pl.n[2];   // [z][y][kx_l]
    const size_t b1 = (size_t)PXl * pl.L[1] * pl.n[2];   // [z][ky][kx_l]
    const size_t blk = (size_t)PXl * pl.n[1] * nzl;      // one (rank, rank) block of the exchanges
    SCB_TRY(ensure_sh_arena(h, (8 * a1 + 4 * b1) * sizeof(C)));
    SCB_TRY(exchange_arenas(h));
    const bool p2p = h->p2p == 1;
    C* base = static_cast<C*>(h->sh_arena);
    C* A1 = base;                 // received x spectra of ALL z, this rank's kx block
    C* S1 = A1 + a1;              // NCCL fallback: F1 output blocked by destination rank
    C* B1 = S1 + a1;              // after F2
    C* C1 = B1 + b1;              // after the z pass, 3 components
    C* DS = C1 + 3 * b1;          // NCCL fallback: B2 output (natural order = blocked by destination rank), 3 components
    C* DR = DS + 3 * a1;          // received for B3: [component][source rank][z_l][y][kx_l]
    const C *twx, *twy, *twz;
    SCB_TRY(get_twiddles<T>(h, pl.L[0], &twx));
    SCB_TRY(get_twiddles<T>(h, pl.L[1], &twy));
    SCB_TRY(get_twiddles<T>(h, pl.L[2], &twz));

    tick(h, 8);
    SCB_NCCL(h, g_nccl.ReduceScatter(rho_partial, slab, slab_elems, nt, ncclSum, h->comm, h->stream));
    tick(h, 14);
    {  // F1 on the z slab; every bin goes straight to the rank that owns its kx block
        XParams<T> p{};
        p.in = slab; p.out = S1; p.tw = twx;
        p.nlines = (long long)pl.n[1] * nzl; p.real_sline = pl.n[0]; p.n_real = pl.n[0]; p.PX = pl.PX; p.scale = (T)1;
        p.split = PXl;
        if (p2p) {
            p.line0 = (long long)me * pl.n[1] * nzl;
            for (int r = 0; r < G; ++r) p.out_peer[r] = static_cast<C*>(h->peer_arena[r]) + (A1 - base);
        } else {
            p.line0 = 0;
            for (int r = 0; r < G; ++r) p.out_peer[r] = S1 + (size_t)r * blk;
        }
        SCB_CUDA(h, launch_x_r2c<T>(pl.L[0], p, 1, h->stream));
    }
    if (p2p) SCB_TRY(rank_barrier(h));
    else SCB_TRY(all_to_all(h, reinterpret_cast<const char*>(S1), reinterpret_cast<char*>(A1), blk * sizeof(C), 1, 0));
    tick(h, 9);
    if (ninl > 0) {  // F2, local
        LinesParams<T> p{};
        p.in = A1; p.out = B1; p.tw = twy;
        p.n_in = pl.n[1]; p.n_out = pl.L[1]; p.ninner = ninl;
        p.in_sline = PXl; p.in_souter = (long long)PXl * pl.n[1];
        p.out_sline = PXl; p.out_souter = (long long)PXl * pl.L[1];
        p.scale = (T)1;
        SCB_CUDA(h, launch_lines<T>(pl.L[1], -1, p, pl.n[2], 1, h->stream));
    }
    tick(h, 10);
    if (ninl > 0) {  // z pass, local, with the single-GPU kernels
        ZParams<T> p{};
        p.in = B1; p.out = C1; p.tw = twz;
        p.out_scomp = (long long)b1;
        p.nz = pl.n[2]; p.ncomp = 3; p.ninner = ninl; p.PX = PXl;
        p.Ly = pl.L[1]; p.Lyg = pl.L[1]; p.ky0 = 0;
        p.kx0 = kx0; p.ninner_g = pl.ninner; p.PXg = pl.PX;
        p.S = static_cast<const T*>(gfree->data); p.S_scomp = gfree->scomp;
        if (gaux) { p.H = static_cast<const C*>(gaux->data); p.H_scomp = gaux->scomp; }
        SCB_TRY(launch_z_pass<T>(h, pl, p, mode, gfree, gaux, B1, C1, b1, 3));
    }
    tick(h, 11);
    if (ninl > 0) {  // B2, local; plane z goes to the owner of its z slab, into the block reserved for this rank
        LinesParams<T> p{};
        p.in = C1; p.tw = twy;
        p.n_in = pl.L[1]; p.n_out = pl.n[1]; p.ninner = ninl;
        p.in_sline = PXl; p.in_souter = (long long)PXl * pl.L[1]; p.in_scomp = (long long)b1;
        p.out_sline = PXl; p.out_souter = (long long)PXl * pl.n[1]; p.out_scomp = (long long)a1;
        p.scale = (T)1;
        if (p2p) {
            p.use_peers = 2;
            p.out_osplit = nzl;
            p.out = DR;
            for (int r = 0; r < G; ++r) p.out_peer[r] = static_cast<C*>(h->peer_arena[r]) + (DR - base) + (size_t)me * blk;
        } else {
            p.out = DS;
        }
        SCB_CUDA(h, launch_lines<T>(pl.L[1], +1, p, pl.n[2], 3, h->stream));
    }
    if (p2p) SCB_TRY(rank_barrier(h));
    else SCB_TRY(all_to_all(h, reinterpret_cast<const char*>(DS), reinterpret_cast<char*>(DR), blk * sizeof(C), 3, a1 * sizeof(C)));
    tick(h, 12);
    const long long ng = (long long)pl.n[0] * pl.n[1] * pl.n[2];
    {  // B3 on the z slab, kx gathered from the ranks' blocks, straight into this rank's slab of the full field
        XParams<T> p{};
        p.in = DR; p.out = efield + (size_t)me * slab_elems; p.tw = twx;
        p.nlines = (long long)pl.n[1] * nzl; p.real_sline = pl.n[0]; p.n_real = pl.n[0]; p.PX = pl.PX;
        p.real_scomp = ng; p.cplx_scomp = (long long)a1;
        p.split = PXl; p.sblock = (long long)blk;
        p.scale = (T)(kFPEI / ((double)pl.L[0] * pl.L[1] * pl.L[2]));
        SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], p, 3, h->stream));
    }
    tick(h, 15);
    if (!defer_gather) {
        SCB_NCCL(h, g_nccl.GroupStart());
        for (int c = 0; c < 3; ++c)
            SCB_NCCL(h, g_nccl.AllGather(efield + c * ng + (size_t)me * slab_elems, efield + c * ng, slab_elems, nt, h->comm, h->stream));
        SCB_NCCL(h, g_nccl.GroupEnd());
    }
    tick(h, 13);
    h->t_pass = true;
    h->t_coll = true;
    h->launches += 5 + 4;
    return SCB_OK;
}

template <typename T>
int run_solve_sharded(scb_handle* h, const T* rho_partial, T* efield, const Plan& pl, const double delta[3], double gamma,
                      int mode, const double offset[3], bool defer_gather = false) {
    if (shard_by_kx(h, pl)) return run_solve_sharded_kx<T>(h, rho_partial, efield, pl, delta, gamma, mode, offset, defer_gather);
    using C = cx_t<T>;
    const int G = h->nranks, me = h->rank;
    const int mdt = sizeof(T) == 8 ? SCB_F64 : SCB_F32;
    const ncclDataType_t nt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    const int nzl = pl.n[2] / G, Lyl = pl.L[1] / G;
    const double zero3[3] = {0, 0, 0};
    const GreenEntry* gfree = nullptr;
    const GreenEntry* gaux = nullptr;
    h->t_green = false;
    SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, zero3, mdt, 0), &gfree));
    if (mode == 1) {
        SCB_TRY(get_green(h, pl, make_key(pl, delta, gamma, offset, mdt, 1), &gaux));
        for (auto& e : h->green)
            if (e.key == make_key(pl, delta, gamma, zero3, mdt, 0)) gfree = &e;
    }
    const size_t slab_elems = (size_t)pl.n[0] * pl.n[1] * nzl;
    if (h->slab_bytes < slab_elems * sizeof(T)) {
        if (h->slab) { SCB_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(h->slab); h->slab = nullptr; h->slab_bytes = 0; }
        if (cudaMalloc(&h->slab, slab_elems * sizeof(T)) != cudaSuccess) { (void)cudaGetLastError(); return fail(h, SCB_ERR_ALLOC, "slab allocation failed"); }
        h->slab_bytes = slab_elems * sizeof(T);
    }
    T* slab = static_cast<T*>(h->slab);
    const size_t szA = (size_t)pl.PX * pl.n[1] * nzl;       // A_l, D_c
    const size_t szB = (size_t)pl.PX * pl.L[1] * nzl;       // send/recv buffers (= PX*Lyl*nz)
    SCB_TRY(ensure_sh_arena(h, (4 * szA + 8 * szB) * sizeof(C)));
    SCB_TRY(exchange_arenas(h));
    const bool p2p = h->p2p == 1;
    C* A = static_cast<C*>(h->sh_arena);
    C* SB = A + szA;          // F2 output, blocked by destination rank
    C* RB = SB + szB;         // received: [z][ky_l][kx]
    C* Cc = RB + szB;         // 3 components [z][ky_l][kx]
    C* R2 = Cc + 3 * szB;     // 3 components, received: [g][z_l][ky_l][kx]
    C* D = R2 + 3 * szB;      // 3 components [kx][y][z_l]
    const C *twx, *twy, *twz;
    SCB_TRY(get_twiddles<T>(h, pl.L[0], &twx));
    SCB_TRY(get_twiddles<T>(h, pl.L[1], &twy));
    SCB_TRY(get_twiddles<T>(h, pl.L[2], &twz));
    const size_t blk = (size_t)pl.PX * Lyl * nzl;           // elements per (rank, rank) block

    tick(h, 8);
    SCB_NCCL(h, g_nccl.ReduceScatter(rho_partial, slab, slab_elems, nt, ncclSum, h->comm, h->stream));
    tick(h, 14);
    {  // F1 on the slab
        XParams<T> p{};
        p.in = slab; p.out = A; p.tw = twx;
        p.nlines = (long long)pl.n[1] * nzl; p.real_sline = pl.n[0]; p.n_real = pl.n[0]; p.PX = pl.PX; p.scale = (T)1;
        SCB_CUDA(h, launch_x_r2c<T>(pl.L[0], p, 1, h->stream));
    }
    tick(h, 9);
    {  // F2, output blocked by destination rank: [h][z_l][ky_l][kx]
        LinesParams<T> p{};
        p.in = A; p.out = SB; p.tw = twy;
        p.n_in = pl.n[1]; p.n_out = pl.L[1]; p.ninner = pl.ninner;
        p.in_sline = pl.PX; p.in_souter = (long long)pl.PX * pl.n[1];
        p.out_sline = pl.PX; p.out_souter = (long long)pl.PX * Lyl;
        p.out_split = Lyl; p.out_sblock = (long long)blk;
        if (p2p) {  // store block b straight into rank b's RB, at the slot reserved for this rank
            p.use_peers = 1;
            for (int r = 0; r < G; ++r)
                p.out_peer[r] = static_cast<C*>(h->peer_arena[r]) + (RB - A) + (size_t)me * blk;
        }
        p.scale = (T)1;
        SCB_CUDA(h, launch_lines<T>(pl.L[1], -1, p, nzl, 1, h->stream));
    }
    if (p2p) SCB_TRY(rank_barrier(h));
    else SCB_TRY(all_to_all(h, reinterpret_cast<const char*>(SB), reinterpret_cast<char*>(RB), blk * sizeof(C), 1, 0));
    tick(h, 10);
    {  // Z on this rank's ky slab
        ZParams<T> p{};
        p.in = RB; p.out = Cc; p.tw = twz;
        p.out_scomp = (long long)szB;
        p.nz = pl.n[2]; p.ncomp = 3; p.ninner = pl.ninner; p.PX = pl.PX;
        p.Ly = Lyl; p.Lyg = pl.L[1]; p.ky0 = me * Lyl;
        p.S = static_cast<const T*>(gfree->data); p.S_scomp = gfree->scomp;
        if (gaux) { p.H = static_cast<const C*>(gaux->data); p.H_scomp = gaux->scomp; }
        if (p2p) {  // z planes of rank r go straight into rank r's R2, at the slot reserved for this rank
            p.use_peers = 1;
            p.out_split = nzl;
            for (int r = 0; r < G; ++r)
                p.out_peer[r] = static_cast<C*>(h->peer_arena[r]) + (R2 - A) + (size_t)me * blk;
        }
        SCB_CUDA(h, launch_z_fused<T>(pl.L[2], mode == 0 ? GREEN_FREE : GREEN_CATHODE, p, h->stream));
    }
    tick(h, 11);
    if (p2p) SCB_TRY(rank_barrier(h));
    else SCB_TRY(all_to_all(h, reinterpret_cast<const char*>(Cc), reinterpret_cast<char*>(R2), blk * sizeof(C), 3, szB * sizeof(C)));
    {  // B2, input blocked by source rank
        LinesParams<T> p{};
        p.in = R2; p.out = D; p.tw = twy;
        p.n_in = pl.L[1]; p.n_out = pl.n[1]; p.ninner = pl.ninner;
        p.in_sline = pl.PX; p.in_souter = (long long)pl.PX * Lyl;
        p.in_split = Lyl; p.in_sblock = (long long)blk;
        p.out_sline = pl.PX; p.out_souter = (long long)pl.PX * pl.n[1];
        p.in_scomp = (long long)szB; p.out_scomp = (long long)szA;
        p.scale = (T)1;
        SCB_CUDA(h, launch_lines<T>(pl.L[1], +1, p, nzl, 3, h->stream));
    }
    tick(h, 12);
    const long long ng = (long long)pl.n[0] * pl.n[1] * pl.n[2];
    {  // B3 straight into this rank's z slab of the full field
        XParams<T> p{};
        p.in = D; p.out = efield + (size_t)me * slab_elems; p.tw = twx;
        p.nlines = (long long)pl.n[1] * nzl; p.real_sline = pl.n[0]; p.n_real = pl.n[0]; p.PX = pl.PX;
        p.real_scomp = ng; p.cplx_scomp = (long long)szA;
        p.scale = (T)(kFPEI / ((double)pl.L[0] * pl.L[1] * pl.L[2]));
        SCB_CUDA(h, launch_x_c2r<T>(pl.L[0], p, 3, h->stream));
    }
    tick(h, 15);
    if (!defer_gather) {   // scb_step_sharded gathers the slabs itself, overlapped with the interpolation
        SCB_NCCL(h, g_nccl.GroupStart());
        for (int c = 0; c < 3; ++c)
            SCB_NCCL(h, g_nccl.AllGather(efield + c * ng + (size_t)me * slab_elems, efield + c * ng, slab_elems, nt, h->comm, h->stream));
        SCB_NCCL(h, g_nccl.GroupEnd());
    }
    tick(h, 13);
    h->t_pass = true;
    h->t_coll = true;
    h->launches += 5 + 4;  // five pass kernels + four collectives
    return SCB_OK;
}

// Field slabs -> every rank, overlapped with the gather.  The all-gather of E is receive-bound (every rank takes in
// (G-1)/G of the 3*Ng field) and nothing is left to compute on the grid once the last pass has run, so the only work
// that can hide it is the interpolation itself: the slabs are broadcast one by one from the middle of the grid
// outwards (for a bunch the middle planes hold most particles), each is repacked node-major as it lands, and the
// gather runs in passes over the cells whose two z planes have arrived -- after 2, 4, 8, ... slabs -- each pass
// skipping the particles outside its z range (later passes read z alone and fetch x, y for the few they select).
template <typename T>
int gather_field_and_interpolate(scb_handle* h, T* efield, const Plan& pl, int64_t np, const void* x, const void* y, const void* z,
                                 int pdt, const Geom3& g0, void* ex, void* ey, void* ez) {
    const int G = h->nranks;
    const int mdt = sizeof(T) == 8 ? SCB_F64 : SCB_F32;
    const ncclDataType_t nt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    const int nzl = pl.n[2] / G;
    const long long ng = (long long)pl.n[0] * pl.n[1] * pl.n[2];
    const size_t slab_elems = (size_t)pl.n[0] * pl.n[1] * nzl;
    SCB_TRY(ensure_packed(h, (size_t)ng * packed_bytes_per_node(mdt)));
    if (!h->comm_stream) {
        // highest priority: the collectives' few CTAs must get onto the SMs as soon as gather CTAs of an earlier pass retire
        int least = 0, greatest = 0;
        SCB_CUDA(h, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SCB_CUDA(h, cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, greatest));
        SCB_CUDA(h, cudaStreamCreateWithPriority(&h->pack_stream, cudaStreamNonBlocking, greatest));
    }
    if (!h->ev_field) SCB_CUDA(h, cudaEventCreateWithFlags(&h->ev_field, cudaEventDisableTiming));
    for (int i = 0; i < G; ++i) {
        if (!h->ev_bcast[i]) SCB_CUDA(h, cudaEventCreateWithFlags(&h->ev_bcast[i], cudaEventDisableTiming));
        if (!h->ev_pack[i]) SCB_CUDA(h, cudaEventCreateWithFlags(&h->ev_pack[i], cudaEventDisableTiming));
    }
    // middle-out slab order, identical on every rank: G/2-1, G/2, G/2-2, G/2+1, ...
    int order[SCB_MAX_RANKS];
    for (int i = 0, lo = G / 2 - 1, hi = G / 2; i < G;) {
        if (lo >= 0) order[i++] = lo--;
        if (hi < G && i < G) order[i++] = hi++;
    }
    SCB_CUDA(h, cudaEventRecord(h->ev_field, h->stream));          // this rank's slab is complete (B3)
    SCB_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_field, 0));
    for (int i = 0; i < G; ++i) {
        const int sl = order[i];
        SCB_NCCL(h, g_nccl.GroupStart());
        for (int c = 0; c < 3; ++c) {
            T* p = efield + c * ng + (size_t)sl * slab_elems;
            SCB_NCCL(h, g_nccl.Broadcast(p, p, slab_elems, nt, sl, h->comm, h->comm_stream));
        }
        SCB_NCCL(h, g_nccl.GroupEnd());
        SCB_CUDA(h, cudaEventRecord(h->ev_bcast[i], h->comm_stream));
        SCB_CUDA(h, cudaStreamWaitEvent(h->pack_stream, h->ev_bcast[i], 0));
        SCB_CUDA(h, launch_pack_efield(mdt, efield, h->packed, g0, h->pack_stream, (long long)sl * slab_elems, (long long)slab_elems));
        SCB_CUDA(h, cudaEventRecord(h->ev_pack[i], h->pack_stream));
        h->launches += 2;
    }
    // gather passes: after `cnt` slabs the arrived planes form one block [a, b]; its cells are [a*nzl, (b+1)*nzl - 1)
    int done = 0, prev_lo = 0, prev_hi = 0;
    for (int cnt = (G == 2 ? 1 : 2); done < G; cnt = cnt * 2 > G ? G : cnt * 2) {
        int a = G, b = -1;
        for (int i = 0; i < cnt; ++i) { a = order[i] < a ? order[i] : a; b = order[i] > b ? order[i] : b; }
        for (int i = done; i < cnt; ++i) SCB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_pack[i], 0));
        Geom3 g = g0;
        g.zlo = a * nzl;
        g.zhi = (b + 1) * nzl - 1;
        g.exlo = prev_lo;
        g.exhi = prev_hi;
        g.zfirst = done > 0 ? 1 : 0;
        SCB_CUDA(h, launch_interpolate_packed(pdt, mdt, np, x, y, z, h->packed, g, ex, ey, ez, h->stream));
        h->launches += 1;
        prev_lo = g.zlo;
        prev_hi = g.zhi;
        done = cnt;
        if (cnt == G) break;
    }
    return SCB_OK;
}

}  // namespace

extern "C" {

int scb_comm_unique_id(void* uid128) {
    if (!uid128) return SCB_ERR_INVALID_ARG;
    if (!load_nccl()) return SCB_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return SCB_ERR_COMM;
    std::memcpy(uid128, &id, 128);
    return SCB_OK;
}

int scb_comm_init(scb_handle* h, int nranks, int rank, const void* uid128) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!uid128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_comm_init");
    if (nranks > SCB_MAX_RANKS) return fail(h, SCB_ERR_UNSUPPORTED, "at most " + std::to_string(SCB_MAX_RANKS) + " ranks are supported");
    if (!load_nccl()) return fail(h, SCB_ERR_COMM, "libnccl.so.2 could not be loaded");
    SCB_CUDA(h, cudaSetDevice(h->device));
    if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, uid128, 128);
    SCB_NCCL(h, g_nccl.CommInitRank(&h->comm, nranks, id, rank));
    h->nranks = nranks;
    h->rank = rank;
    close_peers(h);          // mappings belong to the previous communicator's ranks
    h->peer_gen = ~0ull;
    h->p2p = -1;
    return SCB_OK;
}

int scb_comm_destroy(scb_handle* h) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (h->comm) {
        cudaStreamSynchronize(h->stream);
        g_nccl.CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->nranks = 1;
    h->rank = 0;
    return SCB_OK;
}

int scb_allreduce_rho(scb_handle* h, void* rho, const int64_t n[3], int mdt) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!h->comm) return fail(h, SCB_ERR_COMM, "scb_comm_init has not been called");
    if (!rho || !valid_dt(mdt)) return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_allreduce_rho");
    SCB_TRY(check_grid(h, n));
    SCB_CUDA(h, cudaSetDevice(h->device));
    SCB_NCCL(h, g_nccl.AllReduce(rho, rho, (size_t)n[0] * n[1] * n[2], mdt == SCB_F64 ? ncclFloat64 : ncclFloat32, ncclSum,
                                 h->comm, h->stream));
    h->launches += 1;
    return SCB_OK;
}

int scb_solve_sharded(scb_handle* h, const void* rho_partial, void* efield, int mdt, const int64_t n[3],
                      const double min_bounds[3], const double max_bounds[3], const double delta[3], double gamma,
                      int at_cathode) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!h->comm) return fail(h, SCB_ERR_COMM, "scb_comm_init has not been called");
    if (!rho_partial || !efield || !delta || !min_bounds || !max_bounds || !valid_dt(mdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_solve_sharded");
    SCB_TRY(check_grid(h, n));
    const Plan pl = make_plan(n);
    if (pl.n[2] % h->nranks != 0 || (pl.PX % h->nranks != 0 && pl.L[1] % h->nranks != 0))
        return fail(h, SCB_ERR_UNSUPPORTED, "nz and the spectrum pitch (or the padded y length) must be multiples of the number of ranks");
    SCB_CUDA(h, cudaSetDevice(h->device));
    double offset[3] = {0.0, 0.0, 0.0};
    if (at_cathode) offset[2] = image_offset_z(mdt, min_bounds[2], max_bounds[2]);
    tick(h, 2);
    int rc;
    if (mdt == SCB_F64)
        rc = run_solve_sharded<double>(h, (const double*)rho_partial, (double*)efield, pl, delta, gamma, at_cathode ? 1 : 0, offset);
    else
        rc = run_solve_sharded<float>(h, (const float*)rho_partial, (float*)efield, pl, delta, gamma, at_cathode ? 1 : 0, offset);
    tick(h, 3);
    h->t_solve = rc == SCB_OK;
    return rc;
}

int scb_step_sharded(scb_handle* h, int64_t np, const void* x, const void* y, const void* z, const void* q, int pdt,
                     void* rho_partial, void* efield, int mdt, const int64_t n[3], const double min_bounds[3],
                     const double max_bounds[3], const double delta[3], double gamma, int at_cathode, void* ex, void* ey,
                     void* ez) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!h->comm) return fail(h, SCB_ERR_COMM, "scb_comm_init has not been called");
    if (np < 0 || !rho_partial || !efield || !delta || !min_bounds || !max_bounds || !valid_dt(mdt) || !valid_dt(pdt))
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_step_sharded");
    if (np > 0 && (!x || !y || !z || !q || !ex || !ey || !ez)) return fail(h, SCB_ERR_INVALID_ARG, "null particle array");
    SCB_TRY(check_grid(h, n));
    const Plan pl = make_plan(n);
    if (pl.n[2] % h->nranks != 0 || (pl.PX % h->nranks != 0 && pl.L[1] % h->nranks != 0))
        return fail(h, SCB_ERR_UNSUPPORTED, "nz and the spectrum pitch (or the padded y length) must be multiples of the number of ranks");
    // Opt-in (SCB_GATHER_OVERLAP=1, read on every call).  Measured on 2 GPUs at config 5: 6.38 ms per step against 5.11 ms
    // for the plain sequence -- one-root broadcasts give up the all-links parallelism of the all-gather, and a filtered
    // gather pass costs almost a full pass (every warp iteration still pays its latency) -- see DESIGN.md section 4.
    const char* ov = std::getenv("SCB_GATHER_OVERLAP");
    const bool overlap = ov && std::atoi(ov) != 0;
    // the slab filter lives in the packed-field gather kernels (default interpolation mode)
    const bool pipelined = overlap && h->nranks > 1 && interp_mode() == 0 && pl.n[2] / h->nranks >= 2;
    SCB_TRY(scb_deposit(h, np, x, y, z, q, pdt, rho_partial, mdt, n, min_bounds, delta, 1));
    if (!pipelined) {
        SCB_TRY(scb_solve_sharded(h, rho_partial, efield, mdt, n, min_bounds, max_bounds, delta, gamma, at_cathode));
        return scb_interpolate(h, np, x, y, z, pdt, efield, mdt, n, min_bounds, delta, ex, ey, ez);
    }
    SCB_CUDA(h, cudaSetDevice(h->device));
    double offset[3] = {0.0, 0.0, 0.0};
    if (at_cathode) offset[2] = image_offset_z(mdt, min_bounds[2], max_bounds[2]);
    tick(h, 2);
    int rc;
    if (mdt == SCB_F64)
        rc = run_solve_sharded<double>(h, (const double*)rho_partial, (double*)efield, pl, delta, gamma, at_cathode ? 1 : 0, offset, true);
    else
        rc = run_solve_sharded<float>(h, (const float*)rho_partial, (float*)efield, pl, delta, gamma, at_cathode ? 1 : 0, offset, true);
    tick(h, 3);
    h->t_solve = true;
    if (rc != SCB_OK) return rc;
    tick(h, 4);
    const Geom3 g = make_geom(n, min_bounds, delta);
    if (mdt == SCB_F64) rc = gather_field_and_interpolate<double>(h, (double*)efield, pl, np, x, y, z, pdt, g, ex, ey, ez);
    else rc = gather_field_and_interpolate<float>(h, (float*)efield, pl, np, x, y, z, pdt, g, ex, ey, ez);
    tick(h, 5);
    h->t_interp = true;
    return rc;
}

}  // extern "C"

// =============================================================================================
// parity hooks for the individual passes (include/spacecharge_b200_debug.h)
#include "../../include/spacecharge_b200_debug.h"

namespace {

template <typename T>
int debug_lines(scb_handle* h, int N, int dir, const void* in, void* out, int n_in, int n_out, int ninner,
                int64_t in_sline, int64_t in_souter, int64_t out_sline, int64_t out_souter, int nouter, double scale) {
    const cx_t<T>* tw;
    SCB_TRY(get_twiddles<T>(h, N, &tw));
    LinesParams<T> p{};
    p.in = static_cast<const cx_t<T>*>(in);
    p.out = static_cast<cx_t<T>*>(out);
    p.tw = tw;
    p.n_in = n_in;
    p.n_out = n_out;
    p.ninner = ninner;
    p.in_sline = in_sline;
    p.in_souter = in_souter;
    p.out_sline = out_sline;
    p.out_souter = out_souter;
    p.scale = (T)scale;
    SCB_CUDA(h, launch_lines<T>(N, dir, p, nouter, 1, h->stream));
    h->launches += 1;
    return SCB_OK;
}

template <typename T>
int debug_x(scb_handle* h, bool r2c, int N, const void* in, void* out, int64_t nlines, int64_t real_sline, int n_real,
            int PX, double scale) {
    const cx_t<T>* tw;
    SCB_TRY(get_twiddles<T>(h, N, &tw));
    XParams<T> p{};
    p.in = in;
    p.out = out;
    p.tw = tw;
    p.nlines = nlines;
    p.real_sline = real_sline;
    p.n_real = n_real;
    p.PX = PX;
    p.scale = (T)scale;
    if (r2c) SCB_CUDA(h, launch_x_r2c<T>(N, p, 1, h->stream));
    else SCB_CUDA(h, launch_x_c2r<T>(N, p, 1, h->stream));
    h->launches += 1;
    return SCB_OK;
}

}  // namespace

extern "C" {

int scb_debug_l2_probe(scb_handle* h, int mode, int64_t buffer_bytes, int iters, double* sector_ops_per_s) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if ((mode != 0 && mode != 1) || buffer_bytes < 4096 || iters < 1 || !sector_ops_per_s)
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_debug_l2_probe");
    SCB_CUDA(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)buffer_bytes / 32 * 32;
    SCB_TRY(ensure_arena(h, bytes));
    SCB_CUDA(h, cudaMemsetAsync(h->arena, 0, bytes, h->stream));
    cudaEvent_t e0, e1;
    SCB_CUDA(h, cudaEventCreate(&e0));
    SCB_CUDA(h, cudaEventCreate(&e1));
    const unsigned grid = 148 * 8;   // 2048 threads per SM, as many as the particle kernels keep resident
    double ops = 0.0;
    float best = 0.f;
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {   // first repetition warms the L2
        cudaEventRecord(e0, h->stream);
        err = launch_probe(mode, h->arena, bytes, iters, grid, h->stream, &ops);
        cudaEventRecord(e1, h->stream);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (err != cudaSuccess) return cuda_fail(h, err, "scb_debug_l2_probe");
    h->launches += 4;
    *sector_ops_per_s = best > 0.f ? ops / (best * 1e-3) : 0.0;
    return SCB_OK;
}

int scb_debug_fft_lines(scb_handle* h, int dt, int N, int dir, const void* in, void* out, int n_in, int n_out,
                        int ninner, int64_t in_sline, int64_t in_souter, int64_t out_sline, int64_t out_souter,
                        int nouter, double scale) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!in || !out || !valid_dt(dt) || !fft_len_supported(N) || nouter < 1 || ninner < 1)
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_debug_fft_lines");
    SCB_CUDA(h, cudaSetDevice(h->device));
    if (dt == SCB_F64)
        return debug_lines<double>(h, N, dir, in, out, n_in, n_out, ninner, in_sline, in_souter, out_sline, out_souter, nouter, scale);
    return debug_lines<float>(h, N, dir, in, out, n_in, n_out, ninner, in_sline, in_souter, out_sline, out_souter, nouter, scale);
}

int scb_debug_fft_x_r2c(scb_handle* h, int dt, int N, const void* in, void* out, int64_t nlines, int64_t real_sline,
                        int n_real, int PX) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!in || !out || !valid_dt(dt) || !fft_len_supported(N) || nlines < 1 || PX < N / 2 + 1)
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_debug_fft_x_r2c");
    SCB_CUDA(h, cudaSetDevice(h->device));
    if (dt == SCB_F64) return debug_x<double>(h, true, N, in, out, nlines, real_sline, n_real, PX, 1.0);
    return debug_x<float>(h, true, N, in, out, nlines, real_sline, n_real, PX, 1.0);
}

int scb_debug_fft_x_c2r(scb_handle* h, int dt, int N, const void* in, void* out, int64_t nlines, int64_t real_sline,
                        int n_real, int PX, double scale) {
    if (!h) return SCB_ERR_INVALID_ARG;
    if (!in || !out || !valid_dt(dt) || !fft_len_supported(N) || nlines < 1 || PX < N / 2 + 1)
        return fail(h, SCB_ERR_INVALID_ARG, "bad argument to scb_debug_fft_x_c2r");
    SCB_CUDA(h, cudaSetDevice(h->device));
    if (dt == SCB_F64) return debug_x<double>(h, false, N, in, out, nlines, real_sline, n_real, PX, scale);
    return debug_x<float>(h, false, N, in, out, nlines, real_sline, n_real, PX, scale);
}

}  // extern "C"
