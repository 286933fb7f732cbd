// sf_gpu.cu -- libstarfish_gpu.so: context management and the C ABI of include/sfgpu.h.
// Build: see starfish_b200/csrc/Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo).
// There is no CPU fallback anywhere in this file: without a CUDA device sfgpu_create() fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <dlfcn.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/sfgpu.h"
#include <cub/device/device_scan.cuh>

#include "sf_fast.cuh"
#include "sf_stream.cuh"
#include "sf_source.cuh"
#include "sf_generic.cuh"
#include "sf_store.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct sfgpu_ctx;
static int fail(sfgpu_ctx *ctx, int code, const char *fmt, ...);

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SFGPU_ENOMEM : SFGPU_ECUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: no link-time dependency, and a host process that already carries an NCCL
// (torch's bundled one) shares it instead of loading a second copy.
// ---------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } sf_ncclUniqueId;
typedef void *sf_ncclComm_t;
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(sf_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(sf_ncclComm_t *, int, sf_ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, sf_ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(sf_ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static const int SF_NCCL_FLOAT64 = 8, SF_NCCL_SUM = 0; // nccl.h: ncclFloat64 = 8, ncclSum = 0

static bool nccl_load(std::string &why)
{
    if (g_nccl.ok) return true;
    const char *names[] = {getenv("SFGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
        return false;
    }
    g_nccl.GetUniqueId = (int (*)(sf_ncclUniqueId *))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(sf_ncclComm_t *, int, sf_ncclUniqueId, int))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, sf_ncclComm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(sf_ncclComm_t))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        why = "libnccl lacks a required symbol";
        return false;
    }
    g_nccl.ok = true;
    return true;
}

// ---------------------------------------------------------------------------------------------
// host-side containers
// ---------------------------------------------------------------------------------------------
struct Records { // growable SoA slab of full particle records on the device
    RecPtrs p{};
    char *slab = nullptr;
    int64_t cap = 0, n = 0;
};

static void rec_bind(Records &r, char *slab, int64_t cap)
{
    double **arr[SF_REC_NDOUBLES] = {&r.p.x, &r.p.y, &r.p.z, &r.p.u, &r.p.v, &r.p.w, &r.p.mpw, &r.p.li, &r.p.lj, &r.p.dt};
    for (int k = 0; k < SF_REC_NDOUBLES; k++) *arr[k] = (double *)(slab + (size_t)k * cap * sizeof(double));
    r.p.tag = (int2 *)(slab + (size_t)SF_REC_NDOUBLES * cap * sizeof(double));
    r.slab = slab;
    r.cap = cap;
}

struct FastStore { // cell-sorted SoA store of the normal particles of one species on one mesh
    FastPtrs p{}, alt{};
    char *slab = nullptr, *alt_slab = nullptr;
    int64_t cap = 0, alt_cap = 0;
    int64_t n = 0;        // physical length (vacant slots included)
    int64_t n_sorted = 0; // prefix covered by the work items of the last sort
    int64_t alive = 0;    // live particles
    bool dirty = false;   // has vacant slots
    int steps_since_sort = 0;
    unsigned *keys = nullptr, *ranks = nullptr, *inv = nullptr; // per particle, sort scratch (key, rank in its cell, inverse permutation)
    int64_t kr_cap = 0;
    // counting pass of the next sort done by the last tiled step (k_fast_step<.., PREP>): keys / ranks / hist describe p[0, keys_n_sorted) except the
    // keys_ndefer slots on the deferred list; valid until anything else edits the store in place
    bool keys_ok = false;
    double keys_pred = 0.0;
    int64_t keys_n_sorted = 0, keys_ndefer = 0;
    unsigned *hist = nullptr, *offs = nullptr;  // per cell key (+1): live particles per key; segment offsets of the sorted prefix
    // streaming step (sf_stream.cuh): histogram accumulated for the next launch, output segment offsets, output cursors
    unsigned *hist_next = nullptr, *offs_out = nullptr, *cursor = nullptr;
    bool stream_ok = false; // hist counts every live particle of p[0,n) and offs describes p[0,n_sorted)
    bool items_ok = false;  // items / d_nitems are the work items of k_fast_step for the current slab (not the streaming kernel's chunks)
    WorkItem *items = nullptr;
    unsigned *defer = nullptr; // slots the tiled kernel leaves to k_fast_deferred
    int64_t defer_cap = 0;
    unsigned *d_nitems = nullptr;
    unsigned max_items = 0, n_items = 0;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int nti = 0, ntj = 0;
    unsigned nkeys = 0;
};

static void fast_bind(FastPtrs &p, char *slab, int64_t cap)
{
    double **arr[7] = {&p.x, &p.y, &p.z, &p.u, &p.v, &p.w, &p.mpw};
    for (int k = 0; k < 7; k++) *arr[k] = (double *)(slab + (size_t)k * cap * sizeof(double));
    p.tag = (int2 *)(slab + (size_t)7 * cap * sizeof(double));
}
#define SF_FAST_BYTES_PER_PARTICLE (7 * sizeof(double) + sizeof(int2))

struct MeshHost {
    MeshDev dev{}; // device pointers inside
    int8_t *bc[4] = {nullptr, nullptr, nullptr, nullptr};
    int *nbr[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *has_seg = nullptr;
    int *seg_offs = nullptr, *seg_ids = nullptr; // node.segments as CSR (sfgpu_mesh_set_segments)
    double4 *seg_xy = nullptr;
    int2 *seg_kind = nullptr;
    double *fields = nullptr; // efi, efj, bfi, bfj packed
    double *node_vol = nullptr;
    bool needs_slow = false; // segments or a CIRCUIT face present
};

struct Pop { // one species on one mesh: MeshData, KM:1314-1427
    FastStore fast;    // normal particles (the bulk)
    Records cur, nxt;  // exceptional particles as full records (this step / next step)
    Records xin, xout; // transfer_particles (being moved / being filled)
    double *dep = nullptr; // packed [SFGPU_NFIELDS][ni][nj] raw per-step deposit
    double *samp = nullptr; // packed running velocity-moment sums (count,u,v,w,uu,vv,ww,mpc), KM:1570-1595
};

struct Species {
    double charge = 0, mass = 0, qm = 0;
    int32_t id_counter = 0; // part_id_counter, KM:80
    std::vector<Pop> pops;  // per mesh
    Records slow;
    double *slow_extra_d = nullptr; // old_x, old_y, old_li, old_lj
    int *slow_extra_i = nullptr;    // bounces, mesh
    int64_t slow_cap = 0, slow_n = 0;
    double sums[5] = {0, 0, 0, 0, 0};
    int64_t n_exited = 0, n_removed = 0, n_absorbed = 0, n_hits = 0;
    int64_t capacity_hint = 0;
    bool step_open = false; // sfgpu_step ran, sfgpu_finish_step pending
    int64_t num_samples = 0; // KM:1557
};

struct sfgpu_ctx {
    int device = 0, domain = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    int64_t launch_total = 0;
    std::vector<MeshHost> meshes;
    MeshDev *d_meshes = nullptr;
    bool meshes_dirty = true;
    std::vector<Species> species;
    StepCounters *d_cnt = nullptr, *h_cnt = nullptr; // device / pinned host
    XferDev *d_xfer = nullptr;
    FastStepArgs *d_args = nullptr; // per-mesh copy of the tiled kernel's arguments for its out-of-line slow path
    char *stage = nullptr; // pinned staging
    size_t stage_bytes = 0;
    double *d_tmp = nullptr; // moments scratch
    size_t tmp_bytes = 0;
    sf_ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    int last_launches = 0;
    bool timing_valid = false;
    int sort_every = 3;      // steps between cell sorts of the fast store (2..5 give the same step time on config B; 3 keeps the kernel on a fresher order)
    int fast_grid[3] = {0, 0, 0}; // CTAs of the tiled kernel (persistent), by halo width
    int fast_halo = 1;       // halo of the accumulation tile the next tiled step runs with (FastGeom): 1 = 15 warps / SM, 2 = 12 warps / SM
    bool fuse_count = false; // SFGPU_FUSE_COUNT=1: the tiled step before a cell sort also does the sort's counting pass (k_fast_step<.., PREP>).  Parity green; measured
                             // on config B: step 0.872 vs 0.887 ms, but that launch costs +0.10-0.17 ms against 0.16 ms for k_sort_count, config C 6.21 vs 6.16 ms: not default
    bool fast_halo_auto = true; // a step that left more than 0.5 % of its deposits to k_fast_deferred switches the context to the wide halo
    int path = 0;            // default step kernel: 0 = tiled in-place step + periodic sort (sf_fast.cuh), 1 = streaming step (sf_stream.cuh)
    int stream_grid = 0;     // CTAs of the streaming kernel (persistent)
    bool hybrid = false;     // tiled path: run the step in which a re-sort is due with the streaming kernel instead of sorting separately
    unsigned *h_cnt2 = nullptr; // pinned: per-mesh work-item counts read back with the step counters
    int last_kernel = 0;     // step kernel of the last sfgpu_step: 0 tiled, 1 streaming, 2 generic
    // the periodic sort of the tiled path keys every particle by the cell of pos + vel*dt*h, i.e. (up to the field kick) the cell it will be deposited
    // into h steps later: the steps between two sorts then see an order that is at most one or two steps away from "grouped by new cell" in EITHER
    // direction (ages 1,0,1 for 3 steps and h = 2) instead of growing stale (ages 1,2,3).  < 0: h = min(2, (sort_every+1)/2); env SFGPU_SORT_PREDICT=h, 0 = current cell
    double sort_predict = -1.0;
    bool sort_gather = true; // cell sort: inverse permutation + gather with coalesced stores (false: one scatter pass; env SFGPU_SORT_GATHER=0)
    bool stream_sort = false; // periodic re-sort of the tiled path: 1 = streaming pass (k_stream_sort), 0 = generic counting sort (equal speed measured)
    bool stream_check = false; // debug: verify the output cursors after every streaming launch
    unsigned long long *d_bad = nullptr;
    Records tmp;             // staging records for download / upload of the fast store
    char *src_tmp = nullptr; // device scratch of sfgpu_source_uniform (sampled particles, flags, ranks, spline, scan storage)
    size_t src_bytes = 0;
    unsigned long long last_fallback = 0, last_flush = 0;
    char *hit_slab = nullptr; // surface-hit list of a step: t,u,v,w,mpw [cap] doubles, seg, mesh [cap] ints, alive [cap] bytes
    size_t hit_cap = 0;
    bool force_sort = false;
    std::string err;
};

static int fail(sfgpu_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

static int rec_reserve(sfgpu_ctx *ctx, Records &r, int64_t need, bool keep)
{
    if (need <= r.cap) return 0;
    int64_t cap = r.cap + r.cap / 2;
    if (cap < need) cap = need;
    if (cap < 1024) cap = 1024;
    cap = (cap + 255) & ~int64_t(255); // keeps every array 2 KiB aligned for 128-bit access
    char *slab = nullptr;
    const size_t bytes = (size_t)cap * (SF_REC_NDOUBLES * sizeof(double) + sizeof(int2));
    CU(cudaMalloc(&slab, bytes));
    Records old = r;
    rec_bind(r, slab, cap);
    if (keep && old.n > 0) {
        const double *src[SF_REC_NDOUBLES] = {old.p.x, old.p.y, old.p.z, old.p.u, old.p.v, old.p.w, old.p.mpw, old.p.li, old.p.lj, old.p.dt};
        double *dst[SF_REC_NDOUBLES] = {r.p.x, r.p.y, r.p.z, r.p.u, r.p.v, r.p.w, r.p.mpw, r.p.li, r.p.lj, r.p.dt};
        for (int k = 0; k < SF_REC_NDOUBLES; k++)
            CU(cudaMemcpyAsync(dst[k], src[k], (size_t)old.n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(r.p.tag, old.p.tag, (size_t)old.n * sizeof(int2), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (old.slab) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(old.slab));
    }
    return 0;
}

static void rec_free(Records &r)
{
    if (r.slab) cudaFree(r.slab);
    r = Records();
}

static int stage_reserve(sfgpu_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->stage_bytes) return 0;
    if (ctx->stage) CU(cudaFreeHost(ctx->stage));
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    size_t want = bytes + bytes / 4;
    CU(cudaMallocHost(&ctx->stage, want));
    ctx->stage_bytes = want;
    return 0;
}


// ---------------------------------------------------------------------------------------------
// fast store
// ---------------------------------------------------------------------------------------------
static int fast_reserve(sfgpu_ctx *ctx, FastStore &f, int64_t need)
{
    if (need <= f.cap) return 0;
    int64_t cap = f.cap + f.cap / 4;
    if (cap < need) cap = need;
    if (cap < 4096) cap = 4096;
    cap = (cap + 255) & ~int64_t(255);
    char *slab = nullptr;
    CU(cudaMalloc(&slab, (size_t)cap * SF_FAST_BYTES_PER_PARTICLE));
    FastPtrs np_{};
    fast_bind(np_, slab, cap);
    if (f.n > 0) {
        const double *src[7] = {f.p.x, f.p.y, f.p.z, f.p.u, f.p.v, f.p.w, f.p.mpw};
        double *dst[7] = {np_.x, np_.y, np_.z, np_.u, np_.v, np_.w, np_.mpw};
        for (int k = 0; k < 7; k++)
            CU(cudaMemcpyAsync(dst[k], src[k], (size_t)f.n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(np_.tag, f.p.tag, (size_t)f.n * sizeof(int2), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (f.slab) CU(cudaFree(f.slab));
    f.slab = slab;
    f.p = np_;
    f.cap = cap;
    return 0;
}

static void fast_free(FastStore &f)
{
    void *ptrs[] = {f.slab, f.alt_slab, f.keys, f.ranks, f.hist, f.offs, f.items, f.d_nitems, f.cub_tmp, f.hist_next, f.offs_out, f.cursor, f.defer, f.inv};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    f = FastStore();
}

static int fast_init_geometry(sfgpu_ctx *ctx, FastStore &f, const MeshDev &m)
{
    f.nti = (m.ni - 1 + SF_TILE - 1) / SF_TILE;
    f.ntj = (m.nj - 1 + SF_TILE - 1) / SF_TILE;
    f.nkeys = (unsigned)f.nti * f.ntj * SF_TILE * SF_TILE;
    CU(cudaMalloc(&f.hist, ((size_t)f.nkeys + 1) * sizeof(unsigned)));
    CU(cudaMalloc(&f.offs, ((size_t)f.nkeys + 1) * sizeof(unsigned)));
    CU(cudaMalloc(&f.hist_next, ((size_t)f.nkeys + 1) * sizeof(unsigned)));
    CU(cudaMalloc(&f.offs_out, ((size_t)f.nkeys + 1) * sizeof(unsigned)));
    CU(cudaMalloc(&f.cursor, ((size_t)f.nkeys + 1) * sizeof(unsigned)));
    CU(cudaMalloc(&f.d_nitems, sizeof(unsigned)));
    CU(cudaMemset(f.d_nitems, 0, sizeof(unsigned)));
    f.cub_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, f.cub_bytes, f.hist, f.offs, (int)(f.nkeys + 1), ctx->stream));
    CU(cudaMalloc(&f.cub_tmp, f.cub_bytes ? f.cub_bytes : 16));
    return 0;
}

// per-particle sort scratch (key, rank, inverse permutation); re-allocating it drops a counting pass a step may have left there
static int fast_reserve_keys(sfgpu_ctx *ctx, FastStore &f)
{
    if (f.kr_cap >= f.n && f.kr_cap > 0) return 0;
    CU(cudaStreamSynchronize(ctx->stream));
    if (f.keys) CU(cudaFree(f.keys));
    if (f.ranks) CU(cudaFree(f.ranks));
    if (f.inv) CU(cudaFree(f.inv));
    f.keys = f.ranks = f.inv = nullptr; f.kr_cap = 0;
    f.keys_ok = false;
    CU(cudaMalloc(&f.keys, (size_t)f.cap * sizeof(unsigned)));
    CU(cudaMalloc(&f.ranks, (size_t)f.cap * sizeof(unsigned)));
    CU(cudaMalloc(&f.inv, (size_t)f.cap * sizeof(unsigned)));
    f.kr_cap = f.cap;
    return 0;
}

// steps ahead the sort key looks (see k_sort_count)
static double sort_horizon(const sfgpu_ctx *ctx) { return ctx->sort_predict >= 0 ? ctx->sort_predict : std::min(2.0, 0.5 * (ctx->sort_every + 1)); }

// K3: counting sort of the fast store by cell key (tile-major) + compaction of vacant slots, out of place;
// rebuilds the work items of the tiled kernel.  sortParticlesToCells (KM:1150-1179) is the closest reference
// member: order only, no result changes beyond summation order.
static int fast_sort(sfgpu_ctx *ctx, int mesh_id, FastStore &f, double dt_pred = 0.0)
{
    f.steps_since_sort = 0;
    if (f.n == 0) {
        f.keys_ok = false;
        f.n_sorted = 0; f.n_items = 0; f.dirty = false;
        CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
        CU(cudaMemsetAsync(f.hist, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
        CU(cudaMemsetAsync(f.offs, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
        f.stream_ok = true;
        return 0;
    }
    if ((uint64_t)f.n >= 0xfffffff0ull) return fail(ctx, SFGPU_EINVAL, "more than 2^32 particles of one species on one GPU mesh are not supported");
    if (f.alt_cap < f.cap) {
        if (f.alt_slab) CU(cudaFree(f.alt_slab));
        f.alt_slab = nullptr; f.alt_cap = 0;
        CU(cudaMalloc(&f.alt_slab, (size_t)f.cap * SF_FAST_BYTES_PER_PARTICLE));
        f.alt_cap = f.cap;
        fast_bind(f.alt, f.alt_slab, f.alt_cap);
    }
    {
        const int rc = fast_reserve_keys(ctx, f);
        if (rc) return rc;
    }
    const unsigned want_items = (unsigned)(f.n / SF_ITEM_MAX + (int64_t)f.nti * f.ntj + 1);
    if (f.max_items < want_items) { // (the streaming step re-sizes the list for its own, smaller chunks)
        if (f.items) CU(cudaFree(f.items));
        f.items = nullptr; f.max_items = 0;
        CU(cudaMalloc(&f.items, (size_t)want_items * sizeof(WorkItem)));
        f.max_items = want_items;
    }
    const unsigned grid = (unsigned)((f.n + 255) / 256);
    const double pred = dt_pred * sort_horizon(ctx);
    if (f.keys_ok && ctx->sort_gather && pred == f.keys_pred && pred != 0 && f.keys_n_sorted <= f.n) {
        // the last tiled step keyed and counted the particles it finished itself; what is left: its deferred particles and the unsorted tail
        const unsigned long long rest = (unsigned long long)f.keys_ndefer + (unsigned long long)(f.n - f.keys_n_sorted);
        if (rest) {
            k_sort_count_fix<<<(unsigned)((rest + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, f.p, f.defer, (unsigned long long)f.keys_ndefer,
                                                                                   (unsigned long long)f.keys_n_sorted, (unsigned long long)f.n, f.ntj, f.hist, f.keys, f.ranks, pred);
            CU(cudaGetLastError());
        } else { // (the launch count below assumes a counting pass)
            ctx->launch_total--;
            ctx->last_launches--;
        }
    } else {
        CU(cudaMemsetAsync(f.hist, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
        k_sort_count<<<(unsigned)((f.n + 256 * SF_COUNT_ILP - 1) / (256 * SF_COUNT_ILP)), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, f.p, (unsigned long long)f.n, f.ntj, f.hist, f.keys, f.ranks, pred);
        CU(cudaGetLastError());
    }
    f.keys_ok = false;
    CU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, f.cub_bytes, f.hist, f.offs, (int)(f.nkeys + 1), ctx->stream));
    if (ctx->sort_gather) { // inverse permutation first (the keys array is reused for it after the fact: ranks -> inv), then a gather with coalesced stores
        k_sort_invert<<<(unsigned)((f.n + 256 * SF_SORT_ILP - 1) / (256 * SF_SORT_ILP)), 256, 0, ctx->stream>>>((unsigned long long)f.n, f.offs, f.keys, f.ranks, f.inv);
        CU(cudaGetLastError());
        const unsigned ggrid = (unsigned)((f.alive + 256 * SF_SORT_ILP - 1) / (256 * SF_SORT_ILP));
        if (ggrid) k_sort_gather<<<ggrid, 256, 0, ctx->stream>>>(f.p, f.alt, (unsigned long long)f.alive, f.inv);
        ctx->launch_total++;
        ctx->last_launches++;
    } else {
        k_sort_scatter<<<grid, 256, 0, ctx->stream>>>(f.p, f.alt, (unsigned long long)f.n, f.offs, f.keys, f.ranks);
    }
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
    const int n_tiles = f.nti * f.ntj;
    k_build_items<<<(n_tiles + 127) / 128, 128, 0, ctx->stream>>>(f.offs, n_tiles, f.items, f.d_nitems, f.max_items);
    CU(cudaGetLastError());
    ctx->launch_total += 4; // count, scan (cub: counted as one), scatter, build_items
    ctx->last_launches += 4;
    unsigned tot[2] = {0, 0};
    CU(cudaMemcpyAsync(&tot[0], f.offs + f.nkeys, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&tot[1], f.d_nitems, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if ((int64_t)tot[0] != f.alive)
        return fail(ctx, SFGPU_ESTATE, "internal: sort kept %u particles, store tracked %lld", tot[0], (long long)f.alive);
    std::swap(f.p, f.alt);
    std::swap(f.slab, f.alt_slab);
    std::swap(f.cap, f.alt_cap);
    f.n = f.n_sorted = (int64_t)tot[0];
    f.n_items = tot[1] < f.max_items ? tot[1] : f.max_items;
    f.dirty = false;
    f.stream_ok = true; // hist = live particles per key, offs = their segment offsets
    f.items_ok = true;
    return 0;
}

// live particles per cell key of a store whose slots were edited in place (tiled steps, uploads): the layout (offs, n_sorted)
// stands, only the histogram the streaming step lays its output out by is rebuilt
static int fast_rehist(sfgpu_ctx *ctx, int mesh_id, FastStore &f)
{
    CU(cudaMemsetAsync(f.hist, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
    if (f.n > 0) {
        k_stream_hist<<<(unsigned)((f.n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, f.p, 0ULL, (unsigned long long)f.n, f.ntj, f.hist);
        CU(cudaGetLastError());
        ctx->launch_total++;
        ctx->last_launches++;
    }
    f.stream_ok = true;
    return 0;
}

static int fast_reserve_alt(sfgpu_ctx *ctx, FastStore &f);

// K3 for a store that is still roughly in cell order: histogram of the current cells, exclusive scan, one streaming pass
// (k_stream_sort) that writes every particle into its cell's segment of the second slab.  Falls back to the generic counting
// sort for a store that was never sorted or whose unsorted tail is large.
static int fast_resort(sfgpu_ctx *ctx, int mesh_id, FastStore &f, double dt_pred = 0.0)
{
    if (!ctx->stream_sort || f.n == 0 || f.n_sorted == 0 || (f.n - f.n_sorted) * 16 > f.n) return fast_sort(ctx, mesh_id, f, dt_pred);
    if ((uint64_t)f.n >= 0xfffffff0ull) return fail(ctx, SFGPU_EINVAL, "more than 2^32 particles of one species on one GPU mesh are not supported");
    int rc = fast_rehist(ctx, mesh_id, f);
    if (rc) return rc;
    rc = fast_reserve_alt(ctx, f);
    if (rc) return rc;
    const unsigned want_items = (unsigned)(f.n / SFR_CHUNK + (int64_t)f.nti * f.ntj + 16);
    if (f.max_items < want_items) {
        if (f.items) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(f.items)); }
        f.items = nullptr; f.max_items = 0;
        CU(cudaMalloc(&f.items, (size_t)want_items * sizeof(WorkItem)));
        f.max_items = want_items;
    }
    const int n_tiles = f.nti * f.ntj;
    CU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, f.cub_bytes, f.hist, f.offs_out, (int)(f.nkeys + 1), ctx->stream));
    CU(cudaMemcpyAsync(f.cursor, f.offs_out, (size_t)f.nkeys * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
    CU(cudaMemsetAsync(f.items, 0, (size_t)f.max_items * sizeof(WorkItem), ctx->stream));
    k_build_chunks<<<(n_tiles + 127) / 128, 128, 0, ctx->stream>>>(f.offs, n_tiles, f.items, f.d_nitems, f.max_items, SFR_CHUNK);
    CU(cudaGetLastError());
    if (f.n > f.n_sorted) {
        const int64_t n_tail = (f.n - f.n_sorted + SFR_CHUNK - 1) / SFR_CHUNK;
        k_build_tail<<<(unsigned)((n_tail + 127) / 128), 128, 0, ctx->stream>>>((unsigned long long)f.n_sorted, (unsigned long long)(f.n - f.n_sorted), f.items, f.d_nitems, f.max_items, SFR_CHUNK);
        CU(cudaGetLastError());
    }
    k_stream_sort<<<want_items, SFR_THREADS, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, f.p, f.alt, f.items, f.max_items, f.cursor, f.ntj);
    CU(cudaGetLastError());
    if (ctx->stream_check) {
        CU(cudaMemsetAsync(ctx->d_bad, 0, sizeof(unsigned long long), ctx->stream));
        k_stream_check<<<(f.nkeys + 255) / 256, 256, 0, ctx->stream>>>(f.offs_out, f.cursor, f.nkeys, ctx->d_bad);
        unsigned long long bad = 0;
        CU(cudaMemcpyAsync(&bad, ctx->d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (bad) return fail(ctx, SFGPU_ESTATE, "internal: %llu cell segments of the streaming sort were not filled exactly", bad);
    }
    // work items of the tiled kernel for the new layout
    CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
    k_build_items<<<(n_tiles + 127) / 128, 128, 0, ctx->stream>>>(f.offs_out, n_tiles, f.items, f.d_nitems, f.max_items);
    CU(cudaGetLastError());
    ctx->launch_total += 6; // hist (counted there), scan, chunks, tail, sort, items
    ctx->last_launches += 5;
    unsigned n_items = 0;
    CU(cudaMemcpyAsync(&n_items, f.d_nitems, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::swap(f.p, f.alt);
    std::swap(f.slab, f.alt_slab);
    std::swap(f.cap, f.alt_cap);
    std::swap(f.offs, f.offs_out);
    f.n = f.n_sorted = f.alive;
    f.n_items = n_items < f.max_items ? n_items : f.max_items;
    f.dirty = false;
    f.steps_since_sort = 0;
    f.stream_ok = true;
    f.items_ok = true;
    return 0;
}

// second slab of the fast store: sort target / output of the streaming step
static int fast_reserve_alt(sfgpu_ctx *ctx, FastStore &f)
{
    if (f.alt_cap >= f.cap) return 0;
    if (f.alt_slab) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(f.alt_slab));
    }
    f.alt_slab = nullptr; f.alt_cap = 0;
    CU(cudaMalloc(&f.alt_slab, (size_t)f.cap * SF_FAST_BYTES_PER_PARTICLE));
    f.alt_cap = f.cap;
    fast_bind(f.alt, f.alt_slab, f.alt_cap);
    return 0;
}

template <bool SEG, int HALO, bool PREP>
static void launch_fast_step_t(sfgpu_ctx *ctx, const FastStepArgs &a, const FastStepArgs *ga)
{
    typedef FastGeom<HALO> G;
    k_fast_step<SEG, HALO, PREP><<<ctx->fast_grid[HALO], G::WARPS * 32, G::WARPS * G::WARP_BYTES, ctx->stream>>>(a, ga);
}

static void launch_fast_step(sfgpu_ctx *ctx, const FastStepArgs &a, const FastStepArgs *ga, int halo, bool seg, bool prep)
{
    const int k = (halo == 1 ? 0 : 4) + (seg ? 2 : 0) + (prep ? 1 : 0);
    switch (k) {
    case 0: launch_fast_step_t<false, 1, false>(ctx, a, ga); break;
    case 1: launch_fast_step_t<false, 1, true>(ctx, a, ga); break;
    case 2: launch_fast_step_t<true, 1, false>(ctx, a, ga); break;
    case 3: launch_fast_step_t<true, 1, true>(ctx, a, ga); break;
    case 4: launch_fast_step_t<false, 2, false>(ctx, a, ga); break;
    case 5: launch_fast_step_t<false, 2, true>(ctx, a, ga); break;
    case 6: launch_fast_step_t<true, 2, false>(ctx, a, ga); break;
    default: launch_fast_step_t<true, 2, true>(ctx, a, ga); break;
    }
}

template <bool SEG, int HALO, bool PREP>
static cudaError_t fast_step_smem_attr()
{
    return cudaFuncSetAttribute(k_fast_step<SEG, HALO, PREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, FastGeom<HALO>::WARPS * FastGeom<HALO>::WARP_BYTES);
}

static FastStepArgs fast_args(sfgpu_ctx *ctx, Species &s, int m, double dt, const SlowPtrs &slow)
{
    Pop &pop = s.pops[m];
    FastStepArgs a{};
    a.m = ctx->meshes[m].dev; a.meshes = ctx->d_meshes; a.mesh_id = m; a.qm = s.qm; a.charge = s.charge; a.dt = dt;
    a.fs = pop.fast.p; a.items = pop.fast.items; a.n_items = pop.fast.d_nitems; a.ntj = pop.fast.ntj;
    a.exc = pop.nxt.p; a.exc_cap = (unsigned long long)pop.nxt.cap;
    a.xfer = ctx->d_xfer; a.slow = slow; a.dep = pop.dep; a.c = ctx->d_cnt;
    a.defer = pop.fast.defer; a.defer_cap = (unsigned)(pop.fast.defer_cap > 0x7fffffff ? 0x7fffffff : pop.fast.defer_cap);
    return a;
}

// caller buffers from sfgpu_host_alloc (or any page-locked memory) take direct DMA copies, everything else is staged
static bool is_pinned(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

static int sync_meshes(sfgpu_ctx *ctx)
{
    if (!ctx->meshes_dirty) return 0;
    std::vector<MeshDev> h(ctx->meshes.size());
    for (size_t k = 0; k < h.size(); k++) h[k] = ctx->meshes[k].dev;
    if (!ctx->d_meshes) CU(cudaMalloc(&ctx->d_meshes, sizeof(MeshDev) * SF_MAX_MESHES));
    CU(cudaMemcpyAsync(ctx->d_meshes, h.data(), sizeof(MeshDev) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // h goes out of scope
    ctx->meshes_dirty = false;
    return 0;
}

#define CHECK_CTX()                                                        \
    do {                                                                   \
        if (!ctx) return fail(nullptr, SFGPU_EINVAL, "null context");      \
        cudaError_t e0_ = cudaSetDevice(ctx->device);                      \
        if (e0_ != cudaSuccess) return fail(ctx, SFGPU_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e0_)); \
    } while (0)
#define CHECK_SP()                                                                                         \
    do {                                                                                                   \
        if (sp < 0 || sp >= (int)ctx->species.size()) return fail(ctx, SFGPU_EINVAL, "bad species id %d", sp); \
    } while (0)
#define CHECK_MESH()                                                                                       \
    do {                                                                                                   \
        if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) return fail(ctx, SFGPU_EINVAL, "bad mesh id %d", mesh_id); \
    } while (0)

static inline unsigned grid_for(int64_t n, int block, int max_blocks)
{
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (unsigned)g;
}

// ---------------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_abi_version(void) { return SFGPU_ABI_VERSION; }

extern "C" const char *sfgpu_last_error(sfgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

extern "C" int sfgpu_create(int device, int domain_type, sfgpu_ctx **out)
{
    if (!out) return fail(nullptr, SFGPU_EINVAL, "out is null");
    *out = nullptr;
    if (domain_type < SFGPU_XY || domain_type > SFGPU_ZR) return fail(nullptr, SFGPU_EINVAL, "bad domain type %d", domain_type);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SFGPU_ECUDA, "no CUDA device (%s): libstarfish_gpu has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, SFGPU_EINVAL, "device %d out of range [0,%d)", device, ndev);
    sfgpu_ctx *ctx = new (std::nothrow) sfgpu_ctx();
    if (!ctx) return fail(nullptr, SFGPU_ENOMEM, "host allocation failed");
    ctx->device = device;
    ctx->domain = domain_type;
    int rc = [&]() -> int {
        CU(cudaSetDevice(device));
        CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&ctx->ev0));
        CU(cudaEventCreate(&ctx->ev1));
        CU(cudaEventCreate(&ctx->evk0));
        CU(cudaEventCreate(&ctx->evk1));
        CU(cudaEventCreate(&ctx->evt0));
        CU(cudaEventCreate(&ctx->evt1));
        CU(cudaMalloc(&ctx->d_cnt, sizeof(StepCounters)));
        CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(StepCounters), ctx->stream));
        CU(cudaMallocHost(&ctx->h_cnt, sizeof(StepCounters)));
        CU(cudaMallocHost(&ctx->h_cnt2, sizeof(unsigned) * SF_MAX_MESHES));
        if (const char *e = getenv("SFGPU_HYBRID")) ctx->hybrid = atoi(e) != 0;
        if (const char *e = getenv("SFGPU_STREAM_SORT")) ctx->stream_sort = atoi(e) != 0;
        CU(cudaMalloc(&ctx->d_xfer, sizeof(XferDev) * SF_MAX_MESHES));
        CU(cudaMalloc(&ctx->d_args, sizeof(FastStepArgs) * SF_MAX_MESHES));
        if (const char *e = getenv("SFGPU_SORT_GATHER")) ctx->sort_gather = atoi(e) != 0;
        if (const char *e = getenv("SFGPU_SORT_PREDICT")) ctx->sort_predict = atof(e);
        if (const char *e = getenv("SFGPU_FUSE_COUNT")) ctx->fuse_count = atoi(e) != 0;
        if (const char *e = getenv("SFGPU_SORT_EVERY")) ctx->sort_every = atoi(e) > 0 ? atoi(e) : ctx->sort_every;
        CU((fast_step_smem_attr<false, 1, false>())); CU((fast_step_smem_attr<false, 1, true>()));
        CU((fast_step_smem_attr<true, 1, false>())); CU((fast_step_smem_attr<true, 1, true>()));
        CU((fast_step_smem_attr<false, 2, false>())); CU((fast_step_smem_attr<false, 2, true>()));
        CU((fast_step_smem_attr<true, 2, false>())); CU((fast_step_smem_attr<true, 2, true>()));
        int nsm = 0, per_sm[3] = {0, 0, 0};
        CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], k_fast_step<false, 1, false>, FastGeom<1>::WARPS * 32, FastGeom<1>::WARPS * FastGeom<1>::WARP_BYTES));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[2], k_fast_step<false, 2, false>, FastGeom<2>::WARPS * 32, FastGeom<2>::WARPS * FastGeom<2>::WARP_BYTES));
        if (per_sm[1] < 1 || per_sm[2] < 1) return fail(ctx, SFGPU_ECUDA, "k_fast_step does not fit on this device");
        for (int h = 1; h <= 2; h++) {
            ctx->fast_grid[h] = nsm * per_sm[h];
            if (const char *e = getenv("SFGPU_FAST_GRID")) ctx->fast_grid[h] = atoi(e) > 0 ? atoi(e) : ctx->fast_grid[h]; // occupancy experiments
        }
        if (const char *e = getenv("SFGPU_HALO")) { // 1 | 2: fixed; anything else: start narrow, widen when deposits miss the tile
            if (atoi(e) == 1 || atoi(e) == 2) { ctx->fast_halo = atoi(e); ctx->fast_halo_auto = false; }
        }
        CU(cudaFuncSetAttribute(k_stream_step, cudaFuncAttributeMaxDynamicSharedMemorySize, SFS_SMEM_BYTES));
        int per_sm_s = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_s, k_stream_step, SFS_THREADS, SFS_SMEM_BYTES));
        if (per_sm_s < 1) return fail(ctx, SFGPU_ECUDA, "k_stream_step does not fit on this device");
        ctx->stream_grid = nsm * per_sm_s;
        if (const char *e = getenv("SFGPU_STREAM_GRID")) ctx->stream_grid = atoi(e) > 0 ? atoi(e) : ctx->stream_grid;
        if (const char *e = getenv("SFGPU_PATH")) ctx->path = strcmp(e, "stream") == 0 ? 1 : 0;
        ctx->stream_check = getenv("SFGPU_STREAM_CHECK") != nullptr;
        CU(cudaMalloc(&ctx->d_bad, sizeof(unsigned long long)));
        if (getenv("SFGPU_DEBUG")) fprintf(stderr, "sfgpu: k_stream_step %d CTAs/SM x %d threads, %d B dynamic smem per CTA, grid %d, path %s\n", per_sm_s, SFS_THREADS, (int)SFS_SMEM_BYTES, ctx->stream_grid, ctx->path ? "stream" : "tiled");
        if (getenv("SFGPU_DEBUG"))
            fprintf(stderr, "sfgpu: k_fast_step halo 1: %d CTAs/SM x %d warps, %d B dynamic smem per CTA, grid %d; halo 2: %d CTAs/SM x %d warps, %d B, grid %d; halo %d%s\n",
                    per_sm[1], FastGeom<1>::WARPS, FastGeom<1>::WARPS * FastGeom<1>::WARP_BYTES, ctx->fast_grid[1], per_sm[2], FastGeom<2>::WARPS,
                    FastGeom<2>::WARPS * FastGeom<2>::WARP_BYTES, ctx->fast_grid[2], ctx->fast_halo, ctx->fast_halo_auto ? " (auto)" : "");
        // bit-parity self test: a*b+c must round twice
        double *d = nullptr, h = 0;
        CU(cudaMalloc(&d, sizeof(double)));
        const double a = 1.0 + ldexp(1.0, -30), b = 1.0 - ldexp(1.0, -30), c = -1.0;
        k_selftest_fmad<<<1, 1, 0, ctx->stream>>>(a, b, c, d);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(d));
        if (h != 0.0) // fused: -2^-60; unfused: a*b rounds to 1.0 -> 0
            return fail(ctx, SFGPU_ESTATE, "library was built with FMA contraction enabled (self test gave %g); rebuild with -fmad=false", h);
        return 0;
    }();
    if (rc) {
        g_last_error = ctx->err;
        sfgpu_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return 0;
}

extern "C" void sfgpu_destroy(sfgpu_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
    for (auto &s : ctx->species) {
        for (auto &p : s.pops) {
            rec_free(p.cur); rec_free(p.nxt); rec_free(p.xin); rec_free(p.xout);
            fast_free(p.fast);
            if (p.dep) cudaFree(p.dep);
            if (p.samp) cudaFree(p.samp);
        }
        rec_free(s.slow);
        if (s.slow_extra_d) cudaFree(s.slow_extra_d);
        if (s.slow_extra_i) cudaFree(s.slow_extra_i);
    }
    for (auto &m : ctx->meshes) {
        for (int f = 0; f < 4; f++) {
            if (m.bc[f]) cudaFree(m.bc[f]);
            if (m.nbr[f]) cudaFree(m.nbr[f]);
        }
        if (m.has_seg) cudaFree(m.has_seg);
        if (m.seg_offs) cudaFree(m.seg_offs);
        if (m.seg_ids) cudaFree(m.seg_ids);
        if (m.seg_xy) cudaFree(m.seg_xy);
        if (m.seg_kind) cudaFree(m.seg_kind);
        if (m.fields) cudaFree(m.fields);
        if (m.node_vol) cudaFree(m.node_vol);
    }
    rec_free(ctx->tmp);
    if (ctx->d_meshes) cudaFree(ctx->d_meshes);
    if (ctx->hit_slab) cudaFree(ctx->hit_slab);
    if (ctx->d_cnt) cudaFree(ctx->d_cnt);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    if (ctx->d_xfer) cudaFree(ctx->d_xfer);
    if (ctx->d_args) cudaFree(ctx->d_args);
    if (ctx->d_bad) cudaFree(ctx->d_bad);
    if (ctx->src_tmp) cudaFree(ctx->src_tmp);
    if (ctx->h_cnt2) cudaFreeHost(ctx->h_cnt2);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->d_tmp) cudaFree(ctx->d_tmp);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evk0) cudaEventDestroy(ctx->evk0);
    if (ctx->evk1) cudaEventDestroy(ctx->evk1);
    if (ctx->evt0) cudaEventDestroy(ctx->evt0);
    if (ctx->evt1) cudaEventDestroy(ctx->evt1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---------------------------------------------------------------------------------------------
// mesh + fields
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_mesh_add(sfgpu_ctx *ctx, int32_t ni, int32_t nj, const double x0[2], const double dh[2],
                              const int8_t *const bc[4], const int32_t *const nbr[4], const uint8_t *has_seg,
                              const double *node_vol, int32_t *mesh_id)
{
    CHECK_CTX();
    if (ni < 2 || nj < 2 || !x0 || !dh || !bc || !mesh_id) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_add: bad arguments");
    if (!(dh[0] > 0) || !(dh[1] > 0)) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_add: spacing must be positive");
    if ((int)ctx->meshes.size() >= SF_MAX_MESHES) return fail(ctx, SFGPU_EINVAL, "at most %d meshes", SF_MAX_MESHES);
    if (!ctx->species.empty()) return fail(ctx, SFGPU_ESTATE, "add all meshes before the first species");
    MeshHost m;
    MeshDev &d = m.dev;
    d.ni = ni; d.nj = nj; d.domain = ctx->domain;
    d.x0 = x0[0]; d.y0 = x0[1]; d.dhx = dh[0]; d.dhy = dh[1];
    // UM:131-135 xd = x0 + (n-1)*dh, KM:700-706 shift by (xd - x0): same two roundings as Java
    const double xd0 = x0[0] + (ni - 1) * dh[0], xd1 = x0[1] + (nj - 1) * dh[1];
    d.lenx = xd0 - x0[0];
    d.leny = xd1 - x0[1];
    d.nim1 = (double)(ni - 1);
    d.njm1 = (double)(nj - 1);
    d.rdhx = 1.0 / dh[0];
    d.rdhy = 1.0 / dh[1];
    d.fastdiv = 1;
    for (int k = 0; k < 2; k++) { // sf_div_exact: normal divisor well inside the exponent range, significand not all ones
        int e = 0;
        const double fr = frexp(dh[k], &e);
        if (e < -200 || e > 200 || fr >= 1.0 - 0x1p-52 || getenv("SFGPU_NO_FASTDIV")) d.fastdiv = 0;
    }
    const size_t plane = (size_t)ni * nj;
    for (int f = 0; f < 4; f++) {
        const int len = (f == SFGPU_FACE_RIGHT || f == SFGPU_FACE_LEFT) ? nj : ni;
        if (!bc[f]) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_add: bc[%d] is null", f);
        CU(cudaMalloc(&m.bc[f], len));
        CU(cudaMemcpy(m.bc[f], bc[f], len, cudaMemcpyHostToDevice));
        d.bc[f] = m.bc[f];
        for (int k = 0; k < len; k++)
            if (bc[f][k] == SFGPU_BC_CIRCUIT) m.needs_slow = true;
        d.nbr[f] = nullptr;
        if (nbr && nbr[f]) {
            CU(cudaMalloc(&m.nbr[f], sizeof(int) * 2 * len));
            CU(cudaMemcpy(m.nbr[f], nbr[f], sizeof(int) * 2 * len, cudaMemcpyHostToDevice));
            d.nbr[f] = m.nbr[f];
        }
    }
    d.any_seg = 0;
    d.has_seg = nullptr;
    if (has_seg) {
        for (size_t k = 0; k < plane; k++)
            if (has_seg[k]) { d.any_seg = 1; break; }
        if (d.any_seg) {
            CU(cudaMalloc(&m.has_seg, plane));
            CU(cudaMemcpy(m.has_seg, has_seg, plane, cudaMemcpyHostToDevice));
            d.has_seg = m.has_seg;
            m.needs_slow = true;
        }
    }
    CU(cudaMalloc(&m.fields, 4 * plane * sizeof(double)));
    CU(cudaMemset(m.fields, 0, 4 * plane * sizeof(double)));
    d.efi = m.fields; d.efj = m.fields + plane; d.bfi = m.fields + 2 * plane; d.bfj = m.fields + 3 * plane;
    d.has_b = 0;
    d.node_vol = nullptr;
    if (node_vol) {
        CU(cudaMalloc(&m.node_vol, plane * sizeof(double)));
        CU(cudaMemcpy(m.node_vol, node_vol, plane * sizeof(double), cudaMemcpyHostToDevice));
        d.node_vol = m.node_vol;
    }
    d.seg_offs = nullptr; d.seg_ids = nullptr; d.seg_xy = nullptr; d.seg_kind = nullptr;
    d.hits = HitList{};
    d.id = (int)ctx->meshes.size();
    ctx->meshes.push_back(m);
    ctx->meshes_dirty = true;
    *mesh_id = (int)ctx->meshes.size() - 1;
    return 0;
}

// the DIRICHLET / SINK linear segments of node.segments (Mesh.setNodeControlVolumes, MESH:1215-1290) with the
// outcome Material.performSurfaceInteraction has for this material on each of them; the segment part of ProcessBoundary
// (KM:482-603) then runs on the device and no particle of this mesh is handed to the host for it.
extern "C" int sfgpu_mesh_set_segments(sfgpu_ctx *ctx, int32_t mesh_id, int32_t n_seg, const double *x1, const double *y1, const double *x2,
                                       const double *y2, const int32_t *kind, const int32_t *sink, const int32_t *node_offs, const int32_t *node_ids)
{
    CHECK_CTX();
    CHECK_MESH();
    MeshHost &m = ctx->meshes[mesh_id];
    if (n_seg < 0 || (n_seg > 0 && (!x1 || !y1 || !x2 || !y2 || !kind || !node_offs))) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_set_segments: null argument");
    const size_t plane = (size_t)m.dev.ni * m.dev.nj;
    void *old[] = {m.seg_offs, m.seg_ids, m.seg_xy, m.seg_kind};
    CU(cudaStreamSynchronize(ctx->stream));
    for (void *q : old)
        if (q) CU(cudaFree(q));
    m.seg_offs = m.seg_ids = nullptr; m.seg_xy = nullptr; m.seg_kind = nullptr;
    m.dev.seg_offs = m.dev.seg_ids = nullptr; m.dev.seg_xy = nullptr; m.dev.seg_kind = nullptr;
    ctx->meshes_dirty = true;
    if (n_seg == 0) return 0;
    const int n_ids = node_offs[plane];
    if (node_offs[0] != 0 || n_ids < 0 || (n_ids > 0 && !node_ids)) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_set_segments: bad node table");
    std::vector<uint8_t> hs(plane, 0);
    for (size_t k = 0; k < plane; k++) {
        if (node_offs[k + 1] < node_offs[k]) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_set_segments: node_offs must not decrease");
        for (int q = node_offs[k]; q < node_offs[k + 1]; q++)
            if (node_ids[q] < 0 || node_ids[q] >= n_seg) return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_set_segments: segment id %d out of range", node_ids[q]);
        hs[k] = node_offs[k + 1] > node_offs[k];
    }
    std::vector<double4> xy(n_seg);
    std::vector<int2> kd(n_seg);
    for (int k = 0; k < n_seg; k++) {
        if (kind[k] < 0 || kind[k] > SFGPU_SURFACE_SPECULAR)
            return fail(ctx, SFGPU_EINVAL, "sfgpu_mesh_set_segments: segment %d has outcome %d (0 removed, 1 unchanged, 2 specular); other surface models stay on the host path", k, (int)kind[k]);
        xy[k] = make_double4(x1[k], y1[k], x2[k], y2[k]);
        kd[k] = make_int2(kind[k], sink ? sink[k] : 0);
    }
    CU(cudaMalloc(&m.seg_offs, sizeof(int) * (plane + 1)));
    CU(cudaMalloc(&m.seg_ids, sizeof(int) * (n_ids > 0 ? n_ids : 1)));
    CU(cudaMalloc(&m.seg_xy, sizeof(double4) * n_seg));
    CU(cudaMalloc(&m.seg_kind, sizeof(int2) * n_seg));
    CU(cudaMemcpy(m.seg_offs, node_offs, sizeof(int) * (plane + 1), cudaMemcpyHostToDevice));
    if (n_ids > 0) CU(cudaMemcpy(m.seg_ids, node_ids, sizeof(int) * n_ids, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.seg_xy, xy.data(), sizeof(double4) * n_seg, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.seg_kind, kd.data(), sizeof(int2) * n_seg, cudaMemcpyHostToDevice));
    // the node table defines which nodes own a segment (KM:508-518)
    if (!m.has_seg) CU(cudaMalloc(&m.has_seg, plane));
    CU(cudaMemcpy(m.has_seg, hs.data(), plane, cudaMemcpyHostToDevice));
    m.dev.has_seg = m.has_seg;
    m.dev.any_seg = n_ids > 0 ? 1 : 0;
    m.dev.seg_offs = m.seg_offs; m.dev.seg_ids = m.seg_ids; m.dev.seg_xy = m.seg_xy; m.dev.seg_kind = m.seg_kind;
    // hit list shared by the meshes of the context
    if (!ctx->hit_cap) {
        ctx->hit_cap = 1 << 20;
        CU(cudaMalloc(&ctx->hit_slab, ctx->hit_cap * (2 * sizeof(int) + 5 * sizeof(double) + 1)));
    }
    char *hp = ctx->hit_slab;
    HitList h{};
    h.t = (double *)hp; h.u = h.t + ctx->hit_cap; h.v = h.u + ctx->hit_cap; h.w = h.v + ctx->hit_cap; h.mpw = h.w + ctx->hit_cap;
    h.seg = (int *)(h.mpw + ctx->hit_cap); h.mesh = h.seg + ctx->hit_cap;
    h.alive = (signed char *)(h.mesh + ctx->hit_cap);
    h.cap = ctx->hit_cap;
    h.n = &ctx->d_cnt->n_hits;
    m.dev.hits = h;
    // CIRCUIT faces still need the host
    m.needs_slow = false;
    for (int f = 0; f < 4; f++) {
        const int len = (f == SFGPU_FACE_RIGHT || f == SFGPU_FACE_LEFT) ? m.dev.nj : m.dev.ni;
        std::vector<int8_t> b(len);
        CU(cudaMemcpy(b.data(), m.bc[f], len, cudaMemcpyDeviceToHost));
        for (int k = 0; k < len; k++)
            if (b[k] == SFGPU_BC_CIRCUIT) m.needs_slow = true;
    }
    return 0;
}

// surface hits of the last sfgpu_step, KM:586-602 (any order): what addSurfaceMomentum / addSurfaceMassDeposit /
// boundary_charge need on the host.  Arrays nullable; *n returns the number of hits of the step (copied: min(n, max)).
extern "C" int sfgpu_take_surface_hits(sfgpu_ctx *ctx, int32_t sp, int64_t max, int32_t *mesh, int32_t *seg, double *t, double *u, double *v,
                                       double *w, double *mpw, int8_t *alive, int64_t *n, int64_t *n_absorbed)
{
    CHECK_CTX();
    CHECK_SP();
    Species &s = ctx->species[sp];
    if (n) *n = s.n_hits;
    if (n_absorbed) *n_absorbed = s.n_absorbed;
    int64_t c = s.n_hits < max ? s.n_hits : max;
    if ((int64_t)ctx->hit_cap < c) c = (int64_t)ctx->hit_cap;
    if (c <= 0 || !ctx->hit_slab) return 0;
    const size_t cap = ctx->hit_cap;
    double *hd = (double *)ctx->hit_slab;
    int *hi = (int *)(hd + 5 * cap);
    signed char *ha = (signed char *)(hi + 2 * cap);
    double *dst[5] = {t, u, v, w, mpw};
    for (int k = 0; k < 5; k++)
        if (dst[k]) CU(cudaMemcpyAsync(dst[k], hd + k * cap, c * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (seg) CU(cudaMemcpyAsync(seg, hi, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (mesh) CU(cudaMemcpyAsync(mesh, hi + cap, c * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (alive) CU(cudaMemcpyAsync(alive, ha, c, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int sfgpu_set_fields(sfgpu_ctx *ctx, int32_t mesh_id, const double *efi, const double *efj,
                                const double *bfi, const double *bfj)
{
    CHECK_CTX();
    CHECK_MESH();
    MeshHost &m = ctx->meshes[mesh_id];
    const size_t plane = (size_t)m.dev.ni * m.dev.nj, bytes = plane * sizeof(double);
    if (!efi || !efj) return fail(ctx, SFGPU_EINVAL, "sfgpu_set_fields: efi/efj are required");
    if ((bfi == nullptr) != (bfj == nullptr)) return fail(ctx, SFGPU_EINVAL, "sfgpu_set_fields: give both bfi and bfj or neither");
    const int nf = bfi ? 4 : 2;
    if (is_pinned(efi) && is_pinned(efj) && (!bfi || (is_pinned(bfi) && is_pinned(bfj)))) {
        const double *src[4] = {efi, efj, bfi, bfj};
        for (int k = 0; k < nf; k++)
            CU(cudaMemcpyAsync(m.fields + k * plane, src[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (!bfi && m.dev.has_b) CU(cudaMemsetAsync(m.fields + 2 * plane, 0, 2 * bytes, ctx->stream));
        const int hb = bfi ? 1 : 0;
        if (hb != m.dev.has_b) {
            m.dev.has_b = hb;
            ctx->meshes_dirty = true;
        }
        return 0; // stream ordered: page-locked sources must stay unchanged until the next sfgpu_step / sfgpu_sync returns (sfgpu.h)
    }
    int rc = stage_reserve(ctx, nf * bytes);
    if (rc) return rc;
    // the stage may still feed an earlier async copy
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->stage, efi, bytes);
    memcpy(ctx->stage + bytes, efj, bytes);
    if (bfi) {
        memcpy(ctx->stage + 2 * bytes, bfi, bytes);
        memcpy(ctx->stage + 3 * bytes, bfj, bytes);
    }
    CU(cudaMemcpyAsync(m.fields, ctx->stage, nf * bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int has_b = bfi ? 1 : 0;
    if (!bfi && m.dev.has_b) CU(cudaMemsetAsync(m.fields + 2 * plane, 0, 2 * bytes, ctx->stream));
    if (has_b != m.dev.has_b) {
        m.dev.has_b = has_b;
        ctx->meshes_dirty = true;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// species
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_species_add(sfgpu_ctx *ctx, double charge, double mass, int64_t capacity_hint, int32_t *sp)
{
    CHECK_CTX();
    if (!sp) return fail(ctx, SFGPU_EINVAL, "sp is null");
    if (ctx->meshes.empty()) return fail(ctx, SFGPU_ESTATE, "add a mesh before the first species");
    if (!(mass > 0)) return fail(ctx, SFGPU_EINVAL, "mass must be positive");
    Species s;
    s.charge = charge;
    s.mass = mass;
    s.qm = charge / mass; // Material.java:711
    s.capacity_hint = capacity_hint;
    s.pops.resize(ctx->meshes.size());
    for (size_t k = 0; k < ctx->meshes.size(); k++) {
        const size_t plane = (size_t)ctx->meshes[k].dev.ni * ctx->meshes[k].dev.nj;
        CU(cudaMalloc(&s.pops[k].dep, SFGPU_NFIELDS * plane * sizeof(double)));
        CU(cudaMemset(s.pops[k].dep, 0, SFGPU_NFIELDS * plane * sizeof(double)));
        CU(cudaMalloc(&s.pops[k].samp, SFGPU_NFIELDS * plane * sizeof(double)));
        CU(cudaMemset(s.pops[k].samp, 0, SFGPU_NFIELDS * plane * sizeof(double)));
        int rc = fast_init_geometry(ctx, s.pops[k].fast, ctx->meshes[k].dev);
        if (rc) return rc;
    }
    ctx->species.push_back(s);
    *sp = (int)ctx->species.size() - 1;
    return 0;
}

static int slow_reserve(sfgpu_ctx *ctx, Species &s, int64_t need)
{
    if (need <= s.slow_cap) return 0;
    if (s.slow_n) return fail(ctx, SFGPU_ESTATE, "slow-path list must be taken before it can grow");
    rec_free(s.slow);
    if (s.slow_extra_d) CU(cudaFree(s.slow_extra_d));
    if (s.slow_extra_i) CU(cudaFree(s.slow_extra_i));
    s.slow_extra_d = nullptr; s.slow_extra_i = nullptr; s.slow_cap = 0;
    int rc = rec_reserve(ctx, s.slow, need, false);
    if (rc) return rc;
    CU(cudaMalloc(&s.slow_extra_d, (size_t)s.slow.cap * 4 * sizeof(double)));
    CU(cudaMalloc(&s.slow_extra_i, (size_t)s.slow.cap * 2 * sizeof(int)));
    s.slow_cap = s.slow.cap;
    return 0;
}

static SlowPtrs slow_ptrs(const Species &s)
{
    SlowPtrs sp{};
    sp.rec = s.slow.p;
    sp.cap = (unsigned long long)s.slow_cap;
    if (s.slow_cap) {
        sp.old_x = s.slow_extra_d; sp.old_y = s.slow_extra_d + s.slow_cap;
        sp.old_li = s.slow_extra_d + 2 * s.slow_cap; sp.old_lj = s.slow_extra_d + 3 * s.slow_cap;
        sp.bounces = s.slow_extra_i; sp.mesh = s.slow_extra_i + s.slow_cap;
    }
    return sp;
}

// copy host SoA -> device records [first, first+n) through the pinned stage, chunked
static int upload_records(sfgpu_ctx *ctx, Records &r, int64_t first, const sfgpu_particles *p, int32_t id_base, bool assign_ids)
{
    const int64_t n = p->n;
    const int64_t chunk = 1 << 20;
    int rc = stage_reserve(ctx, (size_t)chunk * (SF_REC_NDOUBLES * sizeof(double) + sizeof(int2)));
    if (rc) return rc;
    const double *src[SF_REC_NDOUBLES] = {p->x, p->y, p->z, p->u, p->v, p->w, p->mpw, p->li, p->lj, p->dt};
    double *dst[SF_REC_NDOUBLES] = {r.p.x, r.p.y, r.p.z, r.p.u, r.p.v, r.p.w, r.p.mpw, r.p.li, r.p.lj, r.p.dt};
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t c = (n - off < chunk) ? n - off : chunk;
        CU(cudaStreamSynchronize(ctx->stream)); // stage reuse
        for (int k = 0; k < SF_REC_NDOUBLES; k++) {
            double *st = (double *)ctx->stage + (size_t)k * chunk;
            if (src[k]) memcpy(st, src[k] + off, (size_t)c * sizeof(double));
            else memset(st, 0, (size_t)c * sizeof(double));
            CU(cudaMemcpyAsync(dst[k] + first + off, st, (size_t)c * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        int2 *tg = (int2 *)((double *)ctx->stage + (size_t)SF_REC_NDOUBLES * chunk);
        for (int64_t q = 0; q < c; q++) {
            tg[q].x = assign_ids ? (int32_t)(id_base + off + q) : (p->id ? p->id[off + q] : -1);
            tg[q].y = p->born_it ? p->born_it[off + q] : 0;
        }
        CU(cudaMemcpyAsync(r.p.tag + first + off, tg, (size_t)c * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int download_records(sfgpu_ctx *ctx, const RecPtrs &r, int64_t first, const sfgpu_particles *p)
{
    const int64_t n = p->n;
    const int64_t chunk = 1 << 20;
    int rc = stage_reserve(ctx, (size_t)chunk * (SF_REC_NDOUBLES * sizeof(double) + sizeof(int2)));
    if (rc) return rc;
    double *dst[SF_REC_NDOUBLES] = {p->x, p->y, p->z, p->u, p->v, p->w, p->mpw, p->li, p->lj, p->dt};
    const double *src[SF_REC_NDOUBLES] = {r.x, r.y, r.z, r.u, r.v, r.w, r.mpw, r.li, r.lj, r.dt};
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t c = (n - off < chunk) ? n - off : chunk;
        for (int k = 0; k < SF_REC_NDOUBLES; k++)
            if (dst[k])
                CU(cudaMemcpyAsync((double *)ctx->stage + (size_t)k * chunk, src[k] + first + off, (size_t)c * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        int2 *tg = (int2 *)((double *)ctx->stage + (size_t)SF_REC_NDOUBLES * chunk);
        if (p->id || p->born_it)
            CU(cudaMemcpyAsync(tg, r.tag + first + off, (size_t)c * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < SF_REC_NDOUBLES; k++)
            if (dst[k]) memcpy(dst[k] + off, (double *)ctx->stage + (size_t)k * chunk, (size_t)c * sizeof(double));
        for (int64_t q = 0; q < c; q++) {
            if (p->id) p->id[off + q] = tg[q].x;
            if (p->born_it) p->born_it[off + q] = tg[q].y;
        }
    }
    return 0;
}


// addParticle(md, part) over c particles that already sit behind the fast store (state + tags at f.p[f.n .. f.n + c)):
// XtoL, plus-edge clamp, -0.5dt rewind, finite-velocity filter (k_inject_fast); updates the store's bookkeeping
static int inject_fast_run(sfgpu_ctx *ctx, Species &s, int mesh_id, int64_t c, double dt_step, bool rewind, int64_t *added)
{
    Pop &pop = s.pops[mesh_id];
    FastStore &f = pop.fast;
    CU(cudaMemsetAsync(&ctx->d_cnt->n_bad, 0, sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(&ctx->d_cnt->n_exc[mesh_id], 0, sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(&ctx->d_cnt->overflow, 0, sizeof(unsigned long long), ctx->stream));
    k_inject_fast<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, s.qm, dt_step, rewind ? 1 : 0, f.p, (unsigned long long)f.n,
                                                                           (unsigned long long)c, pop.cur.p, (unsigned long long)pop.cur.n,
                                                                           (unsigned long long)pop.cur.cap, ctx->d_cnt, f.stream_ok ? f.hist : nullptr, f.ntj);
    ctx->launch_total++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(StepCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_cnt->overflow) return fail(ctx, SFGPU_EOVERFLOW, "internal: record list overflow during injection");
    const int64_t n_exc = (int64_t)ctx->h_cnt->n_exc[mesh_id], n_bad = (int64_t)ctx->h_cnt->n_bad;
    pop.cur.n += n_exc;
    f.n += c;
    f.alive += c - n_exc - n_bad;
    if (n_exc || n_bad) f.dirty = true;
    *added = c - n_bad;
    return 0;
}

// bulk addParticle into the fast store (the common case: lc == null, KM:760-774): chunks of 1M through the
// pinned stage, k_inject_fast applies XtoL + the -0.5dt rewind; the rare particle the fast store cannot
// represent lands in the record list.
static int inject_fast(sfgpu_ctx *ctx, Species &s, int mesh_id, const sfgpu_particles *p, double dt_step, uint32_t flags, int64_t *n_added)
{
    Pop &pop = s.pops[mesh_id];
    FastStore &f = pop.fast;
    const int64_t chunk = 1 << 20;
    int64_t want = f.n + p->n;
    if (f.n == 0 && s.capacity_hint > want) want = s.capacity_hint;
    int rc = fast_reserve(ctx, f, want);
    if (rc) return rc;
    rc = stage_reserve(ctx, (size_t)chunk * SF_FAST_BYTES_PER_PARTICLE);
    if (rc) return rc;
    const bool assign_ids = p->id == nullptr;
    const double *src[7] = {p->x, p->y, p->z, p->u, p->v, p->w, p->mpw};
    double *dst[7] = {f.p.x, f.p.y, f.p.z, f.p.u, f.p.v, f.p.w, f.p.mpw};
    int64_t added = 0;
    for (int64_t off = 0; off < p->n; off += chunk) {
        const int64_t c = (p->n - off < chunk) ? p->n - off : chunk;
        rc = rec_reserve(ctx, pop.cur, pop.cur.n + c, true);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->stream)); // stage reuse
        for (int k = 0; k < 7; k++) {
            double *st = (double *)ctx->stage + (size_t)k * chunk;
            memcpy(st, src[k] + off, (size_t)c * sizeof(double));
            CU(cudaMemcpyAsync(dst[k] + f.n, st, (size_t)c * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        int2 *tg = (int2 *)((double *)ctx->stage + (size_t)7 * chunk);
        for (int64_t q = 0; q < c; q++) {
            tg[q].x = assign_ids ? (int32_t)(s.id_counter + off + q) : p->id[off + q];
            tg[q].y = p->born_it ? p->born_it[off + q] : 0;
        }
        CU(cudaMemcpyAsync(f.p.tag + f.n, tg, (size_t)c * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
        int64_t a = 0;
        rc = inject_fast_run(ctx, s, mesh_id, c, dt_step, (flags & SFGPU_INJECT_REWIND) != 0, &a);
        if (rc) return rc;
        added += a;
    }
    if (assign_ids) s.id_counter += (int32_t)p->n; // KM:797
    if (n_added) *n_added = added;
    return 0;
}

extern "C" int sfgpu_inject(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, const sfgpu_particles *p, double dt_step,
                            uint32_t flags, int64_t *n_added)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (n_added) *n_added = 0;
    if (!p || p->n < 0) return fail(ctx, SFGPU_EINVAL, "sfgpu_inject: bad particle view");
    if (p->n == 0) return 0;
    if (!p->x || !p->y || !p->z || !p->u || !p->v || !p->w || !p->mpw) return fail(ctx, SFGPU_EINVAL, "sfgpu_inject: pos/vel/mpw arrays are required");
    if ((p->li == nullptr) != (p->lj == nullptr)) return fail(ctx, SFGPU_EINVAL, "sfgpu_inject: give both li and lj or neither");
    if ((flags & SFGPU_INJECT_TRANSFER) && (flags & (SFGPU_INJECT_REWIND | SFGPU_INJECT_DEPOSIT_NOW)))
        return fail(ctx, SFGPU_EINVAL, "sfgpu_inject: TRANSFER excludes REWIND and DEPOSIT_NOW");
    for (size_t k = 0; k < ctx->species.size(); k++) // (injection kernels use the context's step counters too)
        if ((int)k != sp && ctx->species[k].step_open)
            return fail(ctx, SFGPU_ESTATE, "species %d has a deferred step open: call sfgpu_finish_step(%d) before injecting into species %d", (int)k, (int)k, (int)sp);
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    Species &s = ctx->species[sp];
    Pop &pop = s.pops[mesh_id];
    if (!p->li && !p->dt && !(flags & (SFGPU_INJECT_TRANSFER | SFGPU_INJECT_DEPOSIT_NOW)))
        return inject_fast(ctx, s, mesh_id, p, dt_step, flags, n_added);
    Records &r = (flags & SFGPU_INJECT_TRANSFER) ? pop.xout : pop.cur;
    const int64_t first = r.n;
    rc = rec_reserve(ctx, r, first + p->n, true);
    if (rc) return rc;
    const bool assign_ids = p->id == nullptr;
    rc = upload_records(ctx, r, first, p, s.id_counter, assign_ids);
    if (rc) return rc;
    if (assign_ids) s.id_counter += (int32_t)p->n; // KM:797
    CU(cudaMemsetAsync(&ctx->d_cnt->n_bad, 0, sizeof(unsigned long long), ctx->stream));
    const int compute_lc = p->li == nullptr;
    const int rewind = (flags & SFGPU_INJECT_REWIND) ? 1 : 0;
    k_inject<<<(unsigned)((p->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, s.qm, dt_step, compute_lc, rewind, r.p, (unsigned long long)first, (unsigned long long)p->n, ctx->d_cnt);
    ctx->launch_total++;
    CU(cudaGetLastError());
    if (flags & SFGPU_INJECT_DEPOSIT_NOW) {
        k_deposit_records<<<(unsigned)((p->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, r.p, (unsigned long long)first, (unsigned long long)p->n, pop.dep, ctx->d_cnt);
        ctx->launch_total++;
    CU(cudaGetLastError());
    }
    unsigned long long n_bad = 0;
    CU(cudaMemcpyAsync(&n_bad, &ctx->d_cnt->n_bad, sizeof n_bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    int64_t added = p->n;
    if (n_bad) {
        // MeshData.addParticle drops non-finite velocities (KM:1357-1361): rare, compact on the host
        std::vector<double> buf((size_t)p->n * SF_REC_NDOUBLES);
        std::vector<int32_t> ids(p->n), born(p->n);
        sfgpu_particles v{};
        v.n = p->n;
        double **arr[SF_REC_NDOUBLES] = {&v.x, &v.y, &v.z, &v.u, &v.v, &v.w, &v.mpw, &v.li, &v.lj, &v.dt};
        for (int k = 0; k < SF_REC_NDOUBLES; k++) *arr[k] = buf.data() + (size_t)k * p->n;
        v.id = ids.data();
        v.born_it = born.data();
        rc = download_records(ctx, r.p, first, &v);
        if (rc) return rc;
        int64_t w = 0;
        for (int64_t q = 0; q < p->n; q++) {
            if (!(std::isfinite(v.u[q]) && std::isfinite(v.v[q]) && std::isfinite(v.w[q]))) continue;
            for (int k = 0; k < SF_REC_NDOUBLES; k++) (*arr[k])[w] = (*arr[k])[q];
            ids[w] = ids[q];
            born[w] = born[q];
            w++;
        }
        v.n = w;
        added = w;
        if (w) {
            rc = upload_records(ctx, r, first, &v, 0, false);
            if (rc) return rc;
        }
    }
    r.n = first + added;
    if (n_added) *n_added = added;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// UniformSource sampled on the device (sf_source.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_source_uniform(sfgpu_ctx *ctx, int32_t sp, const sfgpu_spline *spl, uint32_t flags, double v_drift, double mpw, int32_t born_it,
                                    int64_t num_mp, double dt_step, uint64_t *rng_state, int64_t *n_added)
{
    CHECK_CTX();
    CHECK_SP();
    if (n_added) *n_added = 0;
    if (!spl || !rng_state) return fail(ctx, SFGPU_EINVAL, "sfgpu_source_uniform: spline and rng_state are required");
    if (spl->n_seg < 1 || !spl->x1 || !spl->y1 || !spl->x2 || !spl->y2 || !spl->nx || !spl->ny || !spl->area || !spl->cum_area)
        return fail(ctx, SFGPU_EINVAL, "sfgpu_source_uniform: incomplete spline");
    if (num_mp < 0 || num_mp >= (1LL << 31)) return fail(ctx, SFGPU_EINVAL, "sfgpu_source_uniform: bad particle count");
    if (num_mp == 0) return 0;
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    Species &s = ctx->species[sp];
    const int nmesh = (int)ctx->meshes.size();
    const size_t n = (size_t)num_mp, ns = (size_t)spl->n_seg;
    // scratch layout: 6 doubles per particle, spline arrays, then mesh_of / flag / rank_all / rank_mesh, then the scan storage
    size_t cub_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned *)nullptr, (unsigned *)nullptr, (int)n, ctx->stream));
    const size_t dbl = 6 * n + 8 * ns + 1, need = dbl * sizeof(double) + 4 * n * sizeof(unsigned) + cub_bytes + 256;
    if (ctx->src_bytes < need) {
        if (ctx->src_tmp) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(ctx->src_tmp)); }
        ctx->src_tmp = nullptr; ctx->src_bytes = 0;
        CU(cudaMalloc(&ctx->src_tmp, need + need / 2));
        ctx->src_bytes = need + need / 2;
    }
    double *d = (double *)ctx->src_tmp;
    double *x = d, *y = x + n, *z = y + n, *u = z + n, *v = u + n, *w = v + n, *sd = w + n;
    int *mesh_of = (int *)(d + dbl);
    unsigned *flag = (unsigned *)mesh_of + n, *rank_all = flag + n, *rank_mesh = rank_all + n;
    void *cub_tmp = (void *)(((uintptr_t)(rank_mesh + n) + 255) & ~(uintptr_t)255);
    SplineDev sdv{};
    sdv.n_seg = spl->n_seg;
    sdv.spline_area = spl->spline_area;
    const double *src[8] = {spl->x1, spl->y1, spl->x2, spl->y2, spl->nx, spl->ny, spl->area, spl->cum_area};
    const double **dst[8] = {&sdv.x1, &sdv.y1, &sdv.x2, &sdv.y2, &sdv.nx, &sdv.ny, &sdv.area, &sdv.cum_area};
    {
        std::vector<double> h(8 * ns + 1);
        size_t off = 0;
        for (int k = 0; k < 8; k++) {
            const size_t cnt = k == 7 ? ns + 1 : ns;
            memcpy(h.data() + off, src[k], cnt * sizeof(double));
            *dst[k] = sd + off;
            off += cnt;
        }
        CU(cudaMemcpyAsync(sd, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // h goes out of scope
    }
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_source_uniform<<<grid, 256, 0, ctx->stream>>>(sdv, ctx->domain, (flags & SFGPU_SOURCE_COLD_BEAM) ? 1 : 0, v_drift, dt_step, (unsigned long long)n, (unsigned long long)*rng_state, ctx->d_meshes, nmesh,
                                                     x, y, z, u, v, w, mesh_of);
    CU(cudaGetLastError());
    k_source_flags<<<grid, 256, 0, ctx->stream>>>(mesh_of, (unsigned long long)n, -1, flag);
    CU(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flag, rank_all, (int)n, ctx->stream));
    unsigned last[2] = {0, 0};
    CU(cudaMemcpyAsync(&last[0], rank_all + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&last[1], flag + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const int64_t accepted = (int64_t)last[0] + last[1];
    ctx->launch_total += 3;
    int64_t added = 0;
    for (int m = 0; m < nmesh && accepted > 0; m++) {
        k_source_flags<<<grid, 256, 0, ctx->stream>>>(mesh_of, (unsigned long long)n, m, flag);
        CU(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flag, rank_mesh, (int)n, ctx->stream));
        CU(cudaMemcpyAsync(&last[0], rank_mesh + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(&last[1], flag + n - 1, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        const int64_t c = (int64_t)last[0] + last[1];
        ctx->launch_total += 2;
        if (c == 0) continue;
        Pop &pop = s.pops[m];
        FastStore &f = pop.fast;
        int64_t want = f.n + c;
        if (f.n == 0 && s.capacity_hint > want) want = s.capacity_hint;
        rc = fast_reserve(ctx, f, want);
        if (rc) return rc;
        rc = rec_reserve(ctx, pop.cur, pop.cur.n + c, true);
        if (rc) return rc;
        k_source_append<<<grid, 256, 0, ctx->stream>>>(mesh_of, (unsigned long long)n, m, rank_mesh, rank_all, x, y, z, u, v, w, mpw, s.id_counter, born_it,
                                                        f.p, (unsigned long long)f.n);
        CU(cudaGetLastError());
        ctx->launch_total++;
        int64_t a = 0;
        rc = inject_fast_run(ctx, s, m, c, dt_step, true, &a);
        if (rc) return rc;
        added += a;
    }
    s.id_counter += (int32_t)accepted; // KM:797: every particle that found a mesh took an id
    *rng_state = sf_java_jump(*rng_state, 2ULL * (unsigned long long)num_mp);
    if (n_added) *n_added = added;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
static int read_counters(sfgpu_ctx *ctx)
{
    CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(StepCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int push_xfer_table(sfgpu_ctx *ctx, Species &s)
{
    XferDev h[SF_MAX_MESHES];
    memset(h, 0, sizeof h);
    for (size_t k = 0; k < s.pops.size(); k++) {
        h[k].rec = s.pops[k].xout.p;
        h[k].cap = (unsigned long long)s.pops[k].xout.cap;
    }
    CU(cudaMemcpyAsync(ctx->d_xfer, h, sizeof(XferDev) * s.pops.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // h is on the stack
    return 0;
}

extern "C" int sfgpu_finish_step(sfgpu_ctx *ctx, int32_t sp);
static int enqueue_finish(sfgpu_ctx *ctx, Species &s);

extern "C" int sfgpu_step(sfgpu_ctx *ctx, int32_t sp, double dt, uint32_t flags)
{
    CHECK_CTX();
    CHECK_SP();
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    Species &s = ctx->species[sp];
    const int nmesh = (int)ctx->meshes.size();
    if (s.slow_n) return fail(ctx, SFGPU_ESTATE, "take the slow-path particles of the previous step first");
    if (s.step_open) return fail(ctx, SFGPU_ESTATE, "previous step was deferred: call sfgpu_finish_step first");
    for (size_t k = 0; k < ctx->species.size(); k++) // the step counters (mover sums included) are per context: one open step at a time
        if (ctx->species[k].step_open) return fail(ctx, SFGPU_ESTATE, "species %d has a deferred step open: call sfgpu_finish_step(%d) first", (int)k, (int)k);
    const bool untiled = (flags & SFGPU_STEP_GENERIC) != 0;
    ctx->last_launches = 0;
    int64_t n_total = 0;
    bool needs_slow = false;
    for (int m = 0; m < nmesh; m++) {
        n_total += s.pops[m].cur.n + s.pops[m].fast.alive;
        needs_slow = needs_slow || ctx->meshes[m].needs_slow;
    }
    if (needs_slow) {
        rc = slow_reserve(ctx, s, n_total > 4096 ? n_total : 4096);
        if (rc) return rc;
    }
    bool multi = nmesh > 1;
    for (int m = 0; m < nmesh; m++) multi = multi || s.pops[m].xout.n > 0;
    if (multi) {
        int64_t xcap = n_total > 65536 ? n_total : 65536;
        for (int m = 0; m < nmesh; m++) {
            rc = rec_reserve(ctx, s.pops[m].xout, s.pops[m].xout.n + xcap, true);
            if (rc) return rc;
        }
        rc = push_xfer_table(ctx, s);
        if (rc) return rc;
    }
    bool stream = !untiled && !(flags & SFGPU_STEP_INPLACE) && (ctx->path == 1 || (flags & SFGPU_STEP_STREAM));
    CU(cudaEventRecord(ctx->ev0, ctx->stream)); // "whole step" of sfgpu_last_step_timing: the periodic sort included
    // K3: cell sort + compaction of the fast store.
    //  * streaming step: the kernel re-sorts as it writes; a separate pass only (re-)establishes its invariant;
    //  * tiled step: needs a re-sort every few steps.  When one is due, that step is run by the streaming kernel instead
    //    (it moves, deposits AND leaves the store re-sorted), so the stand-alone sort only runs for a store that was never sorted.
    if (!untiled && !stream && ctx->hybrid && !(flags & SFGPU_STEP_INPLACE)) {
        bool due = false;
        for (int m = 0; m < nmesh; m++) {
            const FastStore &f = s.pops[m].fast;
            if (f.n == 0 || f.n_sorted == 0 || (f.n - f.n_sorted) * 16 > f.n) continue; // (needs the real sort below)
            if (f.steps_since_sort >= ctx->sort_every || ctx->force_sort) due = true;
        }
        stream = due;
    }
    for (int m = 0; m < nmesh; m++) {
        FastStore &f = s.pops[m].fast;
        if (untiled) { f.stream_ok = false; f.keys_ok = false; continue; }
        if (stream) {
            f.keys_ok = false;
            if (!f.stream_ok) {
                if (f.n > 0 && f.n_sorted > 0 && (f.n - f.n_sorted) * 16 <= f.n) rc = fast_rehist(ctx, m, f); // the layout stands, the cells moved
                else rc = fast_sort(ctx, m, f);
                if (rc) return rc;
            }
            continue;
        }
        f.stream_ok = false; // the in-place step moves particles between cells without telling the histogram
        if (f.n == 0) continue;
        const int64_t tail = f.n - f.n_sorted;
        if (f.n_sorted == 0 || !f.items_ok || f.steps_since_sort >= ctx->sort_every || tail * 16 > f.n || ctx->force_sort) {
            rc = fast_resort(ctx, m, f, dt); // counting sort keyed by the cell each particle is about to move into (streaming re-sort pass: SFGPU_STREAM_SORT=1)
            if (rc) return rc;
            f.stream_ok = false;
        }
        f.keys_ok = false; // this step moves the particles: a counting pass nobody consumed is stale
    }
    for (int m = 0; m < nmesh; m++) {
        Pop &pop = s.pops[m];
        // room for every record survivor plus fast-store particles that turn exceptional in this step
        const int64_t exc_room = pop.fast.alive < (1 << 20) ? pop.fast.alive : (1 << 20) + pop.fast.alive / 8;
        rc = rec_reserve(ctx, pop.nxt, pop.cur.n + exc_room + 1024, false);
        if (rc) return rc;
        // records that are normal particles again move to the fast store's tail
        if (pop.cur.n > 0 || multi) {
            rc = fast_reserve(ctx, pop.fast, pop.fast.n + pop.cur.n + (multi ? n_total : 0));
            if (rc) return rc;
        }
        if (stream) {
            FastStore &f = pop.fast;
            if (f.cap > 0) {
                rc = fast_reserve_alt(ctx, f);
                if (rc) return rc;
            }
            const unsigned want_items = (unsigned)(f.n / SFR_CHUNK + (int64_t)f.nti * f.ntj + 16); // (sized for the smaller chunks of k_stream_sort too)
            if (f.max_items < want_items) {
                if (f.items) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(f.items)); }
                f.items = nullptr; f.max_items = 0;
                CU(cudaMalloc(&f.items, (size_t)want_items * sizeof(WorkItem)));
                f.max_items = want_items;
            }
        }
    }
    CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(StepCounters), ctx->stream));
    {
        // cursors that do not start at zero: transfer lists pre-filled by the host, fast-store tails
        unsigned long long pre[2 * SF_MAX_MESHES] = {0};
        for (int m = 0; m < nmesh; m++) {
            pre[m] = (unsigned long long)s.pops[m].xout.n;
            pre[SF_MAX_MESHES + m] = (unsigned long long)(stream ? s.pops[m].fast.alive : s.pops[m].fast.n); // streaming: behind the re-sorted output
        }
        memcpy(ctx->h_cnt->xfer_n, pre, sizeof(unsigned long long) * SF_MAX_MESHES);
        memcpy(ctx->h_cnt->fast_n, pre + SF_MAX_MESHES, sizeof(unsigned long long) * SF_MAX_MESHES);
        CU(cudaMemcpyAsync(ctx->d_cnt->xfer_n, ctx->h_cnt->xfer_n, sizeof(unsigned long long) * SF_MAX_MESHES, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_cnt->fast_n, ctx->h_cnt->fast_n, sizeof(unsigned long long) * SF_MAX_MESHES, cudaMemcpyHostToDevice, ctx->stream));
    }
    const SlowPtrs slow = slow_ptrs(s);
    for (int m = 0; m < nmesh; m++) {
        const size_t plane = (size_t)ctx->meshes[m].dev.ni * ctx->meshes[m].dev.nj;
        CU(cudaMemsetAsync(s.pops[m].dep, 0, SFGPU_NFIELDS * plane * sizeof(double), ctx->stream));
    }
    // moveParticles(false), KM:126, fused with the deposit
    CU(cudaEventRecord(ctx->evk0, ctx->stream));
    for (int m = 0; m < nmesh; m++) {
        Pop &pop = s.pops[m];
        FastStore &f = pop.fast;
        if (stream) {
            // output layout of this launch = exclusive scan of the live population per cell key (accumulated by the previous
            // launch, injections included); the launch accumulates the next one
            CU(cub::DeviceScan::ExclusiveSum(f.cub_tmp, f.cub_bytes, f.hist, f.offs_out, (int)(f.nkeys + 1), ctx->stream));
            CU(cudaMemcpyAsync(f.cursor, f.offs_out, (size_t)f.nkeys * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream));
            CU(cudaMemsetAsync(f.hist_next, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
            { ctx->last_launches++; ctx->launch_total++; }
            ctx->h_cnt2[m] = 0;
            if (f.n > 0) {
                const unsigned chunk = SFS_CHUNK;
                CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
                CU(cudaMemsetAsync(f.items, 0, (size_t)f.max_items * sizeof(WorkItem), ctx->stream)); // count 0 ends a CTA's round-robin walk
                f.items_ok = false; // the list now holds the streaming kernel's chunks
                if (f.n_sorted > 0) {
                    const int n_tiles = f.nti * f.ntj;
                    k_build_chunks<<<(n_tiles + 127) / 128, 128, 0, ctx->stream>>>(f.offs, n_tiles, f.items, f.d_nitems, f.max_items, chunk);
                    CU(cudaGetLastError());
                    { ctx->last_launches++; ctx->launch_total++; }
                }
                if (f.n > f.n_sorted) {
                    const int64_t n_tail = (f.n - f.n_sorted + chunk - 1) / chunk;
                    k_build_tail<<<(unsigned)((n_tail + 127) / 128), 128, 0, ctx->stream>>>((unsigned long long)f.n_sorted, (unsigned long long)(f.n - f.n_sorted), f.items, f.d_nitems, f.max_items, chunk);
                    CU(cudaGetLastError());
                    { ctx->last_launches++; ctx->launch_total++; }
                }
                StreamArgs sa{};
                sa.b = fast_args(ctx, s, m, dt, slow);
                sa.out = f.alt;
                sa.cursor = f.cursor;
                sa.hist_next = f.hist_next;
                sa.max_items = f.max_items;
                CU(cudaMemcpyAsync(ctx->d_args + m, &sa.b, sizeof sa.b, cudaMemcpyHostToDevice, ctx->stream)); // pageable source: staged before return
                k_stream_step<<<ctx->stream_grid, SFS_THREADS, SFS_SMEM_BYTES, ctx->stream>>>(sa, ctx->d_args + m);
                CU(cudaGetLastError());
                { ctx->last_launches++; ctx->launch_total++; }
                { // a later step may be a tiled one (SFGPU_STEP_INPLACE, or the context default): its work items follow the new layout
                    CU(cudaMemsetAsync(f.d_nitems, 0, sizeof(unsigned), ctx->stream));
                    const int n_tiles = f.nti * f.ntj;
                    k_build_items<<<(n_tiles + 127) / 128, 128, 0, ctx->stream>>>(f.offs_out, n_tiles, f.items, f.d_nitems, f.max_items);
                    CU(cudaGetLastError());
                    CU(cudaMemcpyAsync(&ctx->h_cnt2[m], f.d_nitems, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
                    { ctx->last_launches++; ctx->launch_total++; }
                }
                if (ctx->stream_check) {
                    CU(cudaMemsetAsync(ctx->d_bad, 0, sizeof(unsigned long long), ctx->stream));
                    k_stream_check<<<(f.nkeys + 255) / 256, 256, 0, ctx->stream>>>(f.offs_out, f.cursor, f.nkeys, ctx->d_bad);
                    unsigned long long bad = 0;
                    CU(cudaMemcpyAsync(&bad, ctx->d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                    if (bad) return fail(ctx, SFGPU_ESTATE, "internal: %llu cell segments of the streaming step were not filled exactly", bad);
                }
            }
        } else if (f.n > 0) {
            FastStepArgs a = fast_args(ctx, s, m, dt, slow);
            int64_t tail_first = 0;
            // the tiled kernel carries only the common case (no B field, dt > 0); everything else goes through sf_move() in the tail kernel
            const bool tiled = !untiled && f.n_sorted > 0 && f.n_items > 0 && !a.m.has_b && dt > 0 && f.n_sorted < 0x7fffffff;
            if (tiled) {
                if (f.defer_cap < f.n_sorted) {
                    if (f.defer) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(f.defer)); }
                    f.defer = nullptr; f.defer_cap = 0;
                    CU(cudaMalloc(&f.defer, (size_t)f.cap * sizeof(unsigned)));
                    f.defer_cap = f.cap;
                    a.defer = f.defer; a.defer_cap = (unsigned)(f.defer_cap > 0x7fffffff ? 0x7fffffff : f.defer_cap);
                }
                // the step after this one starts with a cell sort: do that sort's counting pass here, on the state being stored anyway
                const bool prep = ctx->fuse_count && ctx->sort_gather && !ctx->stream_sort && !ctx->hybrid && !multi &&
                                  f.steps_since_sort + 1 >= ctx->sort_every && dt * sort_horizon(ctx) != 0;
                if (prep) {
                    rc = fast_reserve_keys(ctx, f);
                    if (rc) return rc;
                    CU(cudaMemsetAsync(f.hist, 0, ((size_t)f.nkeys + 1) * sizeof(unsigned), ctx->stream));
                    a.sort_keys = f.keys; a.sort_ranks = f.ranks; a.sort_hist = f.hist; a.sort_dt = dt * sort_horizon(ctx);
                }
                launch_fast_step(ctx, a, ctx->d_args + m, ctx->fast_halo, a.m.any_seg != 0, prep);
                if (prep) { f.keys_ok = true; f.keys_pred = a.sort_dt; f.keys_n_sorted = f.n_sorted; f.keys_ndefer = 0; }
                CU(cudaGetLastError());
                k_fast_deferred<<<148 * 4, 256, 0, ctx->stream>>>(a); // boundary crossers, tile misses, removals: a fraction of a percent
                CU(cudaGetLastError());
                { ctx->last_launches += 2; ctx->launch_total += 2; }
                tail_first = f.n_sorted;
            }
            if (f.n > tail_first) {
                k_fast_tail<<<grid_for(f.n - tail_first, 256, 148 * 32), 256, 0, ctx->stream>>>(a, (unsigned long long)tail_first, (unsigned long long)(f.n - tail_first));
                CU(cudaGetLastError());
                { ctx->last_launches++; ctx->launch_total++; }
            }
            f.steps_since_sort++;
        }
        if (pop.cur.n == 0) continue;
        k_generic_step<<<grid_for(pop.cur.n, 256, 148 * 32), 256, 0, ctx->stream>>>(
            ctx->d_meshes, m, s.qm, s.charge, dt, 0, pop.cur.p, (unsigned long long)pop.cur.n, pop.nxt.p,
            &ctx->d_cnt->n_out[m], (unsigned long long)pop.nxt.cap, ctx->d_xfer, slow, pop.dep, ctx->d_cnt, stream ? f.alt : f.p,
            (unsigned long long)(stream ? f.alt_cap : f.cap), stream ? f.hist_next : nullptr, f.ntj);
        CU(cudaGetLastError());
        { ctx->last_launches++; ctx->launch_total++; }
    }
    CU(cudaEventRecord(ctx->evk1, ctx->stream));
    // single mesh, nothing deferred: the cross-GPU sum and the running sums are enqueued behind the step kernels, so the
    // whole step costs ONE host synchronisation (the counters below)
    const bool fused_finish = !multi && !(flags & SFGPU_STEP_DEFER_FINISH);
    if (fused_finish) {
        rc = enqueue_finish(ctx, s);
        if (rc) return rc;
    }
    rc = read_counters(ctx);
    if (rc) return rc;
    // transfer sweeps, KM:131-142
    if (multi) {
        for (int loop = 0; loop < 10; loop++) {
            unsigned long long pending = 0;
            for (int m = 0; m < nmesh; m++) pending += ctx->h_cnt->xfer_n[m];
            if (!pending) break;
            if (ctx->h_cnt->overflow) break;
            // what was filled becomes the input of this sweep; fresh lists collect new hand-offs
            for (int m = 0; m < nmesh; m++) {
                Pop &pop = s.pops[m];
                std::swap(pop.xin, pop.xout);
                pop.xin.n = (int64_t)ctx->h_cnt->xfer_n[m];
                pop.xout.n = 0;
                rc = rec_reserve(ctx, pop.xout, pop.xin.cap, false);
                if (rc) return rc;
                const int64_t have = (int64_t)ctx->h_cnt->n_out[m];
                pop.nxt.n = have;
                rc = rec_reserve(ctx, pop.nxt, have + pop.xin.n, true);
                if (rc) return rc;
            }
            rc = push_xfer_table(ctx, s);
            if (rc) return rc;
            CU(cudaMemsetAsync(ctx->d_cnt->xfer_n, 0, sizeof(unsigned long long) * SF_MAX_MESHES, ctx->stream));
            for (int m = 0; m < nmesh; m++) {
                Pop &pop = s.pops[m];
                if (pop.xin.n == 0) continue;
                k_generic_step<<<grid_for(pop.xin.n, 256, 148 * 32), 256, 0, ctx->stream>>>(
                    ctx->d_meshes, m, s.qm, s.charge, dt, 1, pop.xin.p, (unsigned long long)pop.xin.n, pop.nxt.p,
                    &ctx->d_cnt->n_out[m], (unsigned long long)pop.nxt.cap, ctx->d_xfer, slow, pop.dep, ctx->d_cnt, stream ? pop.fast.alt : pop.fast.p,
                    (unsigned long long)(stream ? pop.fast.alt_cap : pop.fast.cap), stream ? pop.fast.hist_next : nullptr, pop.fast.ntj);
                CU(cudaGetLastError());
                { ctx->last_launches++; ctx->launch_total++; }
                pop.xin.n = 0;
            }
            rc = read_counters(ctx);
            if (rc) return rc;
        }
        // anything still pending stays on the lists for the next step ("Failed to transfer all particles", KM:141)
        for (int m = 0; m < nmesh; m++) s.pops[m].xout.n = (int64_t)ctx->h_cnt->xfer_n[m];
    }
    if (ctx->h_cnt->overflow)
        return fail(ctx, SFGPU_EOVERFLOW, "%llu particles did not fit an internal list (records / slow path / mesh hand-off)", ctx->h_cnt->overflow);
    for (int m = 0; m < nmesh; m++) {
        if ((int64_t)ctx->h_cnt->n_defer[m] > s.pops[m].fast.defer_cap) return fail(ctx, SFGPU_EOVERFLOW, "internal: deferred list overflow");
        s.pops[m].fast.keys_ndefer = (int64_t)ctx->h_cnt->n_defer[m];
    }
    for (int m = 0; m < nmesh; m++) {
        Pop &pop = s.pops[m];
        FastStore &f = pop.fast;
        pop.nxt.n = (int64_t)ctx->h_cnt->n_out[m];
        std::swap(pop.cur, pop.nxt);
        pop.nxt.n = 0;
        int64_t fn = (int64_t)ctx->h_cnt->fast_n[m];
        if (stream) { // the second slab now holds the store: re-sorted live particles, then what the record kernels appended
            const int64_t n_sorted = f.alive;
            std::swap(f.p, f.alt);
            std::swap(f.slab, f.alt_slab);
            std::swap(f.cap, f.alt_cap);
            std::swap(f.hist, f.hist_next);
            std::swap(f.offs, f.offs_out);
            if (fn > f.cap) fn = f.cap;
            if (fn < n_sorted) fn = n_sorted;
            f.n = fn;
            f.n_sorted = n_sorted;
            f.stream_ok = true;
            f.steps_since_sort = 1; // the output is ordered by the cell each particle had BEFORE this push
            f.n_items = ctx->h_cnt2[m] < f.max_items ? ctx->h_cnt2[m] : f.max_items;
            f.items_ok = true;
        } else {
            if (fn > f.cap) fn = f.cap;
            if (fn < f.n) fn = f.n;
            f.n = fn;
        }
        const long long delta = ctx->h_cnt->fast_delta[m];
        f.alive += delta;
        f.dirty = f.alive != f.n;
    }
    s.n_exited = (int64_t)ctx->h_cnt->n_exited;
    s.n_removed = (int64_t)ctx->h_cnt->n_removed;
    s.n_absorbed = (int64_t)ctx->h_cnt->n_absorbed;
    s.n_hits = (int64_t)ctx->h_cnt->n_hits;
    if (s.n_hits > (int64_t)ctx->hit_cap && ctx->hit_cap)
        return fail(ctx, SFGPU_EOVERFLOW, "%lld surface hits in one step, the list holds %zu", (long long)s.n_hits, ctx->hit_cap);
    s.slow_n = (int64_t)ctx->h_cnt->n_slow;
    ctx->last_fallback = ctx->h_cnt->n_fallback;
    if (ctx->fast_halo_auto && ctx->fast_halo == 1) { // deposits further than one cell from the sorted cell: this population wants the wide tile
        int64_t nfast = 0;
        for (int m = 0; m < nmesh; m++) nfast += s.pops[m].fast.n;
        if ((int64_t)ctx->last_fallback * 200 > nfast) ctx->fast_halo = 2;
    }
    // too many particles drifted out of their warp tiles: sort before the next step instead of waiting for the interval
    ctx->force_sort = !untiled && !stream && (int64_t)ctx->last_fallback * 64 > n_total;
    ctx->last_kernel = untiled ? 2 : (stream ? 1 : 0);
    if (fused_finish) {
        ctx->timing_valid = true;
        for (int k = 0; k < 5; k++) s.sums[k] = ctx->h_cnt->sums[k];
        return 0;
    }
    s.step_open = true;
    if (flags & SFGPU_STEP_DEFER_FINISH) return 0;
    return sfgpu_finish_step(ctx, sp);
}

// stream-ordered tail of a step: cross-GPU sum of the deposit and of the mover sums (particles are partitioned,
// the mesh is replicated), then updateSamples (KM:1570-1595: the per-step increments of the running sums are the raw deposit)
static int enqueue_finish(sfgpu_ctx *ctx, Species &s)
{
    const int nmesh = (int)ctx->meshes.size();
    if (ctx->comm) {
        for (int m = 0; m < nmesh; m++) {
            const size_t plane = (size_t)ctx->meshes[m].dev.ni * ctx->meshes[m].dev.nj;
            int r = g_nccl.AllReduce(s.pops[m].dep, s.pops[m].dep, SFGPU_NFIELDS * plane, SF_NCCL_FLOAT64, SF_NCCL_SUM, ctx->comm, ctx->stream);
            if (r) return fail(ctx, SFGPU_ENCCL, "ncclAllReduce(deposit): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        }
        int r = g_nccl.AllReduce(ctx->d_cnt->sums, ctx->d_cnt->sums, 5, SF_NCCL_FLOAT64, SF_NCCL_SUM, ctx->comm, ctx->stream);
        if (r) return fail(ctx, SFGPU_ENCCL, "ncclAllReduce(sums): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    for (int m = 0; m < nmesh; m++) {
        const size_t cnt = (size_t)SFGPU_NFIELDS * ctx->meshes[m].dev.ni * ctx->meshes[m].dev.nj;
        k_accumulate<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(s.pops[m].samp, s.pops[m].dep, cnt);
        ctx->launch_total++;
        CU(cudaGetLastError());
    }
    s.num_samples++;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    return 0;
}

// closes the step opened by sfgpu_step(SFGPU_STEP_DEFER_FINISH) (or a multi-mesh step)
extern "C" int sfgpu_finish_step(sfgpu_ctx *ctx, int32_t sp)
{
    CHECK_CTX();
    CHECK_SP();
    Species &s = ctx->species[sp];
    if (!s.step_open) return fail(ctx, SFGPU_ESTATE, "sfgpu_finish_step without an open step");
    int rc = enqueue_finish(ctx, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_cnt->sums, ctx->d_cnt->sums, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->timing_valid = true;
    for (int k = 0; k < 5; k++) s.sums[k] = ctx->h_cnt->sums[k];
    s.step_open = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------
// device planes -> caller buffers: straight DMA into page-locked destinations, staged copy otherwise
static int download_planes(sfgpu_ctx *ctx, const double *src, size_t plane, int nplanes, double *const *out)
{
    const size_t bytes = plane * sizeof(double);
    bool any = false, all_pinned = true;
    for (int f = 0; f < nplanes; f++)
        if (out[f]) { any = true; all_pinned = all_pinned && is_pinned(out[f]); }
    if (!any) return 0;
    if (all_pinned) {
        for (int f = 0; f < nplanes; f++)
            if (out[f]) CU(cudaMemcpyAsync(out[f], src + (size_t)f * plane, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    int rc = stage_reserve(ctx, nplanes * bytes);
    if (rc) return rc;
    int lo = -1, hi = -1; // contiguous run of requested planes -> one copy
    for (int f = 0; f < nplanes; f++)
        if (out[f]) { if (lo < 0) lo = f; hi = f; }
    CU(cudaMemcpyAsync(ctx->stage + lo * bytes, src + (size_t)lo * plane, (hi - lo + 1) * bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int f = lo; f <= hi; f++)
        if (out[f]) memcpy(out[f], ctx->stage + f * bytes, bytes);
    return 0;
}

extern "C" int sfgpu_get_deposit(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS])
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!out) return fail(ctx, SFGPU_EINVAL, "out is null");
    const size_t plane = (size_t)ctx->meshes[mesh_id].dev.ni * ctx->meshes[mesh_id].dev.nj;
    return download_planes(ctx, ctx->species[sp].pops[mesh_id].dep, plane, SFGPU_NFIELDS, out);
}

extern "C" int sfgpu_get_samples(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS], int64_t *num_samples)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (num_samples) *num_samples = ctx->species[sp].num_samples;
    if (!out) return 0;
    const size_t plane = (size_t)ctx->meshes[mesh_id].dev.ni * ctx->meshes[mesh_id].dev.nj;
    return download_planes(ctx, ctx->species[sp].pops[mesh_id].samp, plane, SFGPU_NFIELDS, out);
}

extern "C" int sfgpu_clear_samples(sfgpu_ctx *ctx, int32_t sp)
{
    CHECK_CTX();
    CHECK_SP();
    Species &s = ctx->species[sp];
    for (size_t m = 0; m < s.pops.size(); m++) {
        const size_t plane = (size_t)ctx->meshes[m].dev.ni * ctx->meshes[m].dev.nj;
        CU(cudaMemsetAsync(s.pops[m].samp, 0, SFGPU_NFIELDS * plane * sizeof(double), ctx->stream));
    }
    s.num_samples = 0;
    return 0;
}

extern "C" int sfgpu_get_moments(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *nd, double *u, double *v, double *w)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    MeshHost &m = ctx->meshes[mesh_id];
    if (!m.node_vol) return fail(ctx, SFGPU_ESTATE, "mesh %d was added without node_vol", mesh_id);
    const size_t plane = (size_t)m.dev.ni * m.dev.nj, bytes = plane * sizeof(double);
    if (ctx->tmp_bytes < 4 * bytes) {
        if (ctx->d_tmp) CU(cudaFree(ctx->d_tmp));
        ctx->d_tmp = nullptr;
        ctx->tmp_bytes = 0;
        CU(cudaMalloc(&ctx->d_tmp, 4 * bytes));
        ctx->tmp_bytes = 4 * bytes;
    }
    k_moments<<<(unsigned)((plane + 255) / 256), 256, 0, ctx->stream>>>(ctx->species[sp].pops[mesh_id].dep, m.node_vol, plane, ctx->d_tmp);
    ctx->launch_total++;
    CU(cudaGetLastError());
    double *const out[4] = {nd, u, v, w};
    return download_planes(ctx, ctx->d_tmp, plane, 4, out);
}

extern "C" int sfgpu_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(nullptr, SFGPU_EINVAL, "out is null");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(nullptr, SFGPU_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return 0;
}

extern "C" void sfgpu_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int sfgpu_get_sums(sfgpu_ctx *ctx, int32_t sp, double sums5[5], int64_t *np_alive, int64_t *n_exited, int64_t *n_slow)
{
    CHECK_CTX();
    CHECK_SP();
    Species &s = ctx->species[sp];
    if (sums5) for (int k = 0; k < 5; k++) sums5[k] = s.sums[k];
    if (np_alive) {
        int64_t n = 0;
        for (auto &p : s.pops) n += p.cur.n + p.fast.alive;
        *np_alive = n;
    }
    if (n_exited) *n_exited = s.n_exited;
    if (n_slow) *n_slow = s.slow_n;
    return 0;
}

extern "C" int sfgpu_np(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t *np)
{
    CHECK_CTX();
    CHECK_SP();
    if (!np) return fail(ctx, SFGPU_EINVAL, "np is null");
    Species &s = ctx->species[sp];
    if (mesh_id < 0) {
        int64_t n = 0;
        for (auto &p : s.pops) n += p.cur.n + p.fast.alive;
        *np = n;
        return 0;
    }
    CHECK_MESH();
    *np = s.pops[mesh_id].cur.n + s.pops[mesh_id].fast.alive;
    return 0;
}

extern "C" int sfgpu_take_slowpath(sfgpu_ctx *ctx, int32_t sp, int64_t max, sfgpu_particles *out, sfgpu_slow_extra *extra, int64_t *n)
{
    CHECK_CTX();
    CHECK_SP();
    if (!out || !n) return fail(ctx, SFGPU_EINVAL, "sfgpu_take_slowpath: out/n are required");
    Species &s = ctx->species[sp];
    if (max < s.slow_n) return fail(ctx, SFGPU_EINVAL, "sfgpu_take_slowpath: room for %lld, list holds %lld", (long long)max, (long long)s.slow_n);
    *n = s.slow_n;
    if (s.slow_n == 0) return 0;
    sfgpu_particles v = *out;
    v.n = s.slow_n;
    int rc = download_records(ctx, s.slow.p, 0, &v);
    if (rc) return rc;
    if (extra) {
        const SlowPtrs spt = slow_ptrs(s);
        const size_t nb = (size_t)s.slow_n;
        if (extra->old_x) CU(cudaMemcpy(extra->old_x, spt.old_x, nb * sizeof(double), cudaMemcpyDeviceToHost));
        if (extra->old_y) CU(cudaMemcpy(extra->old_y, spt.old_y, nb * sizeof(double), cudaMemcpyDeviceToHost));
        if (extra->old_li) CU(cudaMemcpy(extra->old_li, spt.old_li, nb * sizeof(double), cudaMemcpyDeviceToHost));
        if (extra->old_lj) CU(cudaMemcpy(extra->old_lj, spt.old_lj, nb * sizeof(double), cudaMemcpyDeviceToHost));
        if (extra->bounces) CU(cudaMemcpy(extra->bounces, spt.bounces, nb * sizeof(int), cudaMemcpyDeviceToHost));
        if (extra->mesh) CU(cudaMemcpy(extra->mesh, spt.mesh, nb * sizeof(int), cudaMemcpyDeviceToHost));
    }
    s.slow_n = 0;
    return 0;
}

// the particle store seen by iterators / output / restart / collisions: the fast store in device order
// (compacted first, so logical index == slot), then the exceptional records
static int compact_if_dirty(sfgpu_ctx *ctx, int mesh_id, FastStore &f)
{
    if (!f.dirty) return 0;
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    return fast_sort(ctx, mesh_id, f);
}

extern "C" int sfgpu_download(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t first, sfgpu_particles *out)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!out) return fail(ctx, SFGPU_EINVAL, "out is null");
    Pop &pop = ctx->species[sp].pops[mesh_id];
    int rc = compact_if_dirty(ctx, mesh_id, pop.fast);
    if (rc) return rc;
    rc = sync_meshes(ctx);
    if (rc) return rc;
    const int64_t total = pop.fast.n + pop.cur.n;
    if (first < 0 || out->n < 0 || first + out->n > total)
        return fail(ctx, SFGPU_EINVAL, "sfgpu_download: range [%lld,%lld) outside [0,%lld)", (long long)first, (long long)(first + out->n), (long long)total);
    if (out->n == 0) return 0;
    const int64_t chunk = 1 << 20;
    int64_t done = 0;
    // part 1: fast store -> records -> host
    const int64_t f_end = std::min<int64_t>(first + out->n, pop.fast.n);
    for (int64_t q = first; q < f_end; q += chunk) {
        const int64_t c = std::min<int64_t>(chunk, f_end - q);
        rc = rec_reserve(ctx, ctx->tmp, chunk, false);
        if (rc) return rc;
        k_fast_to_records<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, pop.fast.p, (unsigned long long)q, (unsigned long long)c, ctx->tmp.p);
        ctx->launch_total++;
        CU(cudaGetLastError());
        sfgpu_particles v = *out;
        double **arr[SF_REC_NDOUBLES] = {&v.x, &v.y, &v.z, &v.u, &v.v, &v.w, &v.mpw, &v.li, &v.lj, &v.dt};
        for (auto a : arr) if (*a) *a += done;
        if (v.id) v.id += done;
        if (v.born_it) v.born_it += done;
        v.n = c;
        rc = download_records(ctx, ctx->tmp.p, 0, &v);
        if (rc) return rc;
        done += c;
    }
    // part 2: records
    if (done < out->n) {
        const int64_t r_first = first + done - pop.fast.n;
        sfgpu_particles v = *out;
        double **arr[SF_REC_NDOUBLES] = {&v.x, &v.y, &v.z, &v.u, &v.v, &v.w, &v.mpw, &v.li, &v.lj, &v.dt};
        for (auto a : arr) if (*a) *a += done;
        if (v.id) v.id += done;
        if (v.born_it) v.born_it += done;
        v.n = out->n - done;
        rc = download_records(ctx, pop.cur.p, r_first, &v);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int sfgpu_upload(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t first, const sfgpu_particles *in)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!in) return fail(ctx, SFGPU_EINVAL, "in is null");
    Pop &pop = ctx->species[sp].pops[mesh_id];
    if (pop.fast.dirty) return fail(ctx, SFGPU_ESTATE, "sfgpu_upload: indices are only stable right after sfgpu_download / sfgpu_sort");
    const int64_t total = pop.fast.n + pop.cur.n;
    if (first < 0 || in->n < 0 || first + in->n > total)
        return fail(ctx, SFGPU_EINVAL, "sfgpu_upload: range outside the store");
    if (!in->x || !in->y || !in->z || !in->u || !in->v || !in->w || !in->mpw || !in->li || !in->lj || !in->dt || !in->id || !in->born_it)
        return fail(ctx, SFGPU_EINVAL, "sfgpu_upload: every array is required");
    if (in->n == 0) return 0;
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    const int64_t f_end = std::min<int64_t>(first + in->n, pop.fast.n);
    const int64_t n_fast = f_end > first ? f_end - first : 0;
    auto view = [&](int64_t off, int64_t n) {
        sfgpu_particles v = *in;
        double **arr[SF_REC_NDOUBLES] = {&v.x, &v.y, &v.z, &v.u, &v.v, &v.w, &v.mpw, &v.li, &v.lj, &v.dt};
        for (auto a : arr) *a += off;
        v.id += off;
        v.born_it += off;
        v.n = n;
        return v;
    };
    // records part first: the fast part may append records behind it
    if (in->n > n_fast) {
        const sfgpu_particles v = view(n_fast, in->n - n_fast);
        rc = upload_records(ctx, pop.cur, first + n_fast - pop.fast.n, &v, 0, false);
        if (rc) return rc;
    }
    const int64_t chunk = 1 << 20;
    for (int64_t off = 0; off < n_fast; off += chunk) {
        const int64_t c = std::min<int64_t>(chunk, n_fast - off);
        rc = rec_reserve(ctx, ctx->tmp, chunk, false);
        if (rc) return rc;
        rc = rec_reserve(ctx, pop.cur, pop.cur.n + c, true);
        if (rc) return rc;
        const sfgpu_particles v = view(off, c);
        rc = upload_records(ctx, ctx->tmp, 0, &v, 0, false);
        if (rc) return rc;
        CU(cudaMemsetAsync(&ctx->d_cnt->n_exc[mesh_id], 0, sizeof(unsigned long long), ctx->stream));
        k_records_to_fast<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, ctx->tmp.p, (unsigned long long)c, pop.fast.p,
                                                                                   (unsigned long long)(first + off), pop.cur.p, (unsigned long long)pop.cur.n,
                                                                                   (unsigned long long)pop.cur.cap, ctx->d_cnt);
        ctx->launch_total++;
        CU(cudaGetLastError());
        unsigned long long n_exc = 0;
        CU(cudaMemcpyAsync(&n_exc, &ctx->d_cnt->n_exc[mesh_id], sizeof n_exc, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        pop.cur.n += (int64_t)n_exc;
        pop.fast.alive -= (int64_t)n_exc;
        pop.fast.stream_ok = false; // edited in place: cells may have changed
        pop.fast.keys_ok = false;
        if (n_exc) pop.fast.dirty = true;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// restart.bin particle section: KineticMaterial.saveRestartData / loadRestartData, KM:904-1000
// ---------------------------------------------------------------------------------------------
static inline uint64_t host_be64(uint64_t v) { return __builtin_bswap64(v); }

extern "C" int sfgpu_restart_save(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, void *buf, int64_t buf_bytes, int64_t *bytes_needed)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    Species &s = ctx->species[sp];
    Pop &pop = s.pops[mesh_id];
    int rc = compact_if_dirty(ctx, mesh_id, pop.fast);
    if (rc) return rc;
    rc = sync_meshes(ctx);
    if (rc) return rc;
    const int64_t np = pop.fast.n + pop.cur.n;
    const int64_t need = 8 + np * SF_RESTART_RECORD_BYTES;
    if (bytes_needed) *bytes_needed = need;
    if (!buf) return 0; // size query
    if (buf_bytes < need) return fail(ctx, SFGPU_EINVAL, "sfgpu_restart_save: buffer of %lld bytes, %lld needed", (long long)buf_bytes, (long long)need);
    unsigned char *out = (unsigned char *)buf;
    const uint64_t np_be = host_be64((uint64_t)np); // out.writeLong(getNp()), KM:909
    memcpy(out, &np_be, 8);
    out += 8;
    const int64_t chunk = 1 << 20;
    rc = stage_reserve(ctx, (size_t)chunk * SF_RESTART_RECORD_BYTES);
    if (rc) return rc;
    if (ctx->tmp_bytes < (size_t)chunk * SF_RESTART_RECORD_BYTES) {
        if (ctx->d_tmp) CU(cudaFree(ctx->d_tmp));
        ctx->d_tmp = nullptr;
        ctx->tmp_bytes = 0;
        CU(cudaMalloc(&ctx->d_tmp, (size_t)chunk * SF_RESTART_RECORD_BYTES));
        ctx->tmp_bytes = (size_t)chunk * SF_RESTART_RECORD_BYTES;
    }
    for (int64_t q = 0; q < np;) {
        // a chunk never straddles the two stores; stream order = fast store, then the exceptional records
        const bool in_fast = q < pop.fast.n;
        const int64_t c = std::min<int64_t>(chunk, (in_fast ? pop.fast.n : np) - q);
        if (in_fast) {
            rc = rec_reserve(ctx, ctx->tmp, chunk, false);
            if (rc) return rc;
            k_fast_to_records<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_meshes, mesh_id, pop.fast.p, (unsigned long long)q, (unsigned long long)c, ctx->tmp.p);
            CU(cudaGetLastError());
            k_restart_pack<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(ctx->tmp.p, 0ULL, (unsigned long long)c, s.mass, (unsigned long long *)ctx->d_tmp);
        } else {
            k_restart_pack<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(pop.cur.p, (unsigned long long)(q - pop.fast.n), (unsigned long long)c, s.mass, (unsigned long long *)ctx->d_tmp);
        }
        CU(cudaGetLastError());
        ctx->launch_total += in_fast ? 2 : 1;
        CU(cudaMemcpyAsync(ctx->stage, ctx->d_tmp, (size_t)c * SF_RESTART_RECORD_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(out + q * SF_RESTART_RECORD_BYTES, ctx->stage, (size_t)c * SF_RESTART_RECORD_BYTES);
        q += c;
    }
    return 0;
}

extern "C" int sfgpu_restart_load(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, const void *buf, int64_t buf_bytes, double dt_step, int64_t *bytes_used, int64_t *n_loaded)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!buf || buf_bytes < 8) return fail(ctx, SFGPU_EINVAL, "sfgpu_restart_load: truncated stream");
    const unsigned char *in = (const unsigned char *)buf;
    uint64_t np_be;
    memcpy(&np_be, in, 8);
    const int64_t np = (int64_t)host_be64(np_be);
    if (np < 0 || buf_bytes < 8 + np * SF_RESTART_RECORD_BYTES) return fail(ctx, SFGPU_EINVAL, "sfgpu_restart_load: %lld records announced, stream too short", (long long)np);
    in += 8;
    if (bytes_used) *bytes_used = 8 + np * SF_RESTART_RECORD_BYTES;
    int64_t loaded = 0;
    const int64_t chunk = 1 << 18;
    std::vector<double> col((size_t)std::min<int64_t>(chunk, std::max<int64_t>(np, 1)) * 10);
    std::vector<int32_t> born((size_t)std::min<int64_t>(chunk, std::max<int64_t>(np, 1)));
    for (int64_t q = 0; q < np; q += chunk) {
        const int64_t c = std::min<int64_t>(chunk, np - q);
        double *x = col.data(), *y = x + c, *z = y + c, *u = z + c, *v = u + c, *w = v + c, *mpw = w + c, *li = mpw + c, *lj = li + c, *dtp = lj + c;
        for (int64_t k = 0; k < c; k++) {
            const unsigned char *r = in + (q + k) * SF_RESTART_RECORD_BYTES;
            uint64_t d[12];
            memcpy(d, r, SF_RESTART_RECORD_BYTES);
            auto dbl = [&](int i) { const uint64_t b = host_be64(d[i]); double o; memcpy(&o, &b, 8); return o; };
            x[k] = dbl(0); u[k] = dbl(1); y[k] = dbl(2); v[k] = dbl(3); z[k] = dbl(4); w[k] = dbl(5);
            li[k] = dbl(6); lj[k] = dbl(7); dtp[k] = dbl(8); mpw[k] = dbl(9); // d[10] = mass: the species constant
            born[k] = (int32_t)__builtin_bswap32((uint32_t)(d[11] & 0xffffffffu)); // born_it is written first
        }
        // loadRestartData goes through addParticle(md, part) (KM:978): caller-supplied lc, the -0.5dt rewind is applied
        // again and ids are re-assigned from part_id_counter (reference quirk)
        sfgpu_particles pv{};
        pv.n = c; pv.x = x; pv.y = y; pv.z = z; pv.u = u; pv.v = v; pv.w = w; pv.mpw = mpw; pv.li = li; pv.lj = lj; pv.dt = dtp;
        pv.id = nullptr; pv.born_it = born.data();
        int64_t added = 0;
        int rc = sfgpu_inject(ctx, sp, mesh_id, &pv, dt_step, SFGPU_INJECT_REWIND, &added);
        if (rc) return rc;
        loaded += added;
    }
    if (n_loaded) *n_loaded = loaded;
    return 0;
}

extern "C" int sfgpu_sort(sfgpu_ctx *ctx, int32_t sp)
{
    CHECK_CTX();
    CHECK_SP();
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    Species &s = ctx->species[sp];
    for (int m = 0; m < (int)s.pops.size(); m++) {
        rc = fast_sort(ctx, m, s.pops[m].fast);
        if (rc) return rc;
    }
    return 0;
}

// per-cell particle lists for the consumers that bin particles by cell themselves (DSMC.java:194-252, MCC.java:167-216,
// KineticMaterial.sortParticlesToCells KM:1150-1179): cell-sorts the store and returns, for every cell c = i*(nj-1) + j, the index of its first
// particle in sfgpu_download / sfgpu_upload order and its population.  Particles [0, *n_sorted) are covered; the few exceptional records
// (stale lc, residual dt) follow at [*n_sorted, np) unsorted.
extern "C" int sfgpu_cell_lists(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t *cell_first, int32_t *cell_count, int64_t *n_sorted)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!cell_first || !cell_count) return fail(ctx, SFGPU_EINVAL, "sfgpu_cell_lists: null output");
    int rc = sync_meshes(ctx);
    if (rc) return rc;
    FastStore &f = ctx->species[sp].pops[mesh_id].fast;
    rc = fast_sort(ctx, mesh_id, f);
    if (rc) return rc;
    const MeshDev &m = ctx->meshes[mesh_id].dev;
    const int nci = m.ni - 1, ncj = m.nj - 1;
    for (int64_t c = 0; c < (int64_t)nci * ncj; c++) { cell_first[c] = 0; cell_count[c] = 0; }
    if (n_sorted) *n_sorted = f.n_sorted;
    if (f.n == 0) return 0;
    std::vector<unsigned> offs((size_t)f.nkeys + 1);
    CU(cudaMemcpyAsync(offs.data(), f.offs, offs.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (unsigned key = 0; key < f.nkeys; key++) {
        const int tile = (int)(key / (SF_TILE * SF_TILE)), cell = (int)(key % (SF_TILE * SF_TILE));
        const int ci = (tile / f.ntj) * SF_TILE + cell / SF_TILE, cj = (tile % f.ntj) * SF_TILE + cell % SF_TILE;
        if (ci >= nci || cj >= ncj) continue;
        cell_first[(int64_t)ci * ncj + cj] = offs[key];
        cell_count[(int64_t)ci * ncj + cj] = (int32_t)(offs[key + 1] - offs[key]);
    }
    return 0;
}

extern "C" int sfgpu_set_sort_interval(sfgpu_ctx *ctx, int32_t steps)
{
    CHECK_CTX();
    if (steps < 1) return fail(ctx, SFGPU_EINVAL, "sort interval must be >= 1");
    ctx->sort_every = steps;
    return 0;
}

extern "C" int sfgpu_set_tile_halo(sfgpu_ctx *ctx, int32_t halo)
{
    CHECK_CTX();
    if (halo < 0 || halo > 2) return fail(ctx, SFGPU_EINVAL, "tile halo must be 0 (automatic), 1 or 2");
    ctx->fast_halo_auto = halo == 0;
    ctx->fast_halo = halo == 0 ? 1 : halo;
    return 0;
}

extern "C" int sfgpu_get_tile_halo(sfgpu_ctx *ctx, int32_t *halo, int32_t *automatic)
{
    CHECK_CTX();
    if (halo) *halo = ctx->fast_halo;
    if (automatic) *automatic = ctx->fast_halo_auto ? 1 : 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// multi GPU
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_comm_unique_id(void *id128)
{
    if (!id128) return fail(nullptr, SFGPU_EINVAL, "id128 is null");
    std::string why;
    if (!nccl_load(why)) return fail(nullptr, SFGPU_ENCCL, "%s", why.c_str());
    sf_ncclUniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r) return fail(nullptr, SFGPU_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    memcpy(id128, &id, sizeof id);
    return 0;
}

extern "C" int sfgpu_comm_init(sfgpu_ctx *ctx, int32_t nranks, int32_t rank, const void *id128)
{
    CHECK_CTX();
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(ctx, SFGPU_EINVAL, "sfgpu_comm_init: bad arguments");
    std::string why;
    if (!nccl_load(why)) return fail(ctx, SFGPU_ENCCL, "%s", why.c_str());
    if (ctx->comm) return fail(ctx, SFGPU_ESTATE, "communicator already attached");
    sf_ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    int r = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r) {
        ctx->comm = nullptr;
        return fail(ctx, SFGPU_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    ctx->nranks = nranks;
    ctx->rank = rank;
    return 0;
}

extern "C" int sfgpu_deposit_device_ptr(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, void **ptr, int64_t *count)
{
    CHECK_CTX();
    CHECK_SP();
    CHECK_MESH();
    if (!ptr || !count) return fail(ctx, SFGPU_EINVAL, "ptr/count are required");
    *ptr = ctx->species[sp].pops[mesh_id].dep;
    *count = (int64_t)SFGPU_NFIELDS * ctx->meshes[mesh_id].dev.ni * ctx->meshes[mesh_id].dev.nj;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------------------------
extern "C" int sfgpu_last_step_timing(sfgpu_ctx *ctx, float *ms_total, float *ms_kernel, int32_t *launches)
{
    CHECK_CTX();
    if (!ctx->timing_valid) return fail(ctx, SFGPU_ESTATE, "no step has run yet");
    if (ms_total) CU(cudaEventElapsedTime(ms_total, ctx->ev0, ctx->ev1));
    if (ms_kernel) CU(cudaEventElapsedTime(ms_kernel, ctx->evk0, ctx->evk1));
    if (launches) *launches = ctx->last_launches;
    return 0;
}

extern "C" int sfgpu_last_step_kernel(sfgpu_ctx *ctx, int32_t *kind)
{
    CHECK_CTX();
    if (!kind) return fail(ctx, SFGPU_EINVAL, "kind is null");
    *kind = ctx->last_kernel;
    return 0;
}

extern "C" int sfgpu_sync(sfgpu_ctx *ctx)
{
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int sfgpu_timer_start(sfgpu_ctx *ctx)
{
    CHECK_CTX();
    CU(cudaEventRecord(ctx->evt0, ctx->stream));
    return 0;
}

extern "C" int sfgpu_timer_stop(sfgpu_ctx *ctx, float *ms)
{
    CHECK_CTX();
    if (!ms) return fail(ctx, SFGPU_EINVAL, "ms is null");
    CU(cudaEventRecord(ctx->evt1, ctx->stream));
    CU(cudaEventSynchronize(ctx->evt1));
    CU(cudaEventElapsedTime(ms, ctx->evt0, ctx->evt1));
    return 0;
}

extern "C" int sfgpu_launch_count(sfgpu_ctx *ctx, int64_t *n)
{
    CHECK_CTX();
    if (!n) return fail(ctx, SFGPU_EINVAL, "n is null");
    *n = ctx->launch_total;
    return 0;
}

extern "C" int sfgpu_last_step_counters(sfgpu_ctx *ctx, int64_t *n_fallback)
{
    CHECK_CTX();
    if (n_fallback) *n_fallback = (int64_t)ctx->last_fallback;
    return 0;
}

#include "sf_multi.cuh"
