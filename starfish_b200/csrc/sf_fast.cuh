// sf_fast.cuh -- the B200 fast path: fused gather / push / locate / deposit over the cell-sorted fast store.
//
// Layout (DESIGN.md "Data layout"): structure of arrays x,y,z,u,v,w,mpw (+tag), sorted by cell inside 8x8-cell
// tiles.  A warp owns one work item = a run of <= SF_ITEM_MAX particles of one tile.  Per 32-particle batch:
//   1. coalesced loads of the 7 state doubles, lc = XtoL(pos) recomputed (true division, bit exact),
//   2. sf_move(): E/B gather (L1-cached global loads: cell-sorted lanes hit the same 4 nodes), kick, substeps,
//      domain boundaries -- the same code the generic kernel runs, so both round identically,
//   3. in-place store of pos/vel (48 B; unchanged components are skipped),
//   4. deposit: every lane publishes its 4 bilinear weights and 7 moment values in warp-private shared memory;
//      then lane (n,f) walks the 32 particles accumulating weight_n * value_f in a register and adds the sum
//      to the WARP-PRIVATE accumulation tile only when the cell changes (particles are cell-sorted, so runs are
//      long).  The tile is private to the warp: plain read-modify-write, no shared-memory atomics (FP64
//      shared atomics are CAS loops on sm_100a, ~64 cycles per warp instruction).
//   5. after the last batch the tile is added to the global deposit with FP64 REDs (one per touched node
//      and field per work item instead of 29 per particle).
// Particles that drifted out of the tile + halo since the last sort deposit with global REDs (rare).
// Exceptional particles (stale lc, residual dt, mesh hand-off, slow path) leave the fast store for the record
// lists of the generic kernel.
#pragma once
#include "sf_generic.cuh"

#define SF_FAST_WARPS 4 // warps per CTA of the tiled kernel
#define SF_WROW 34      // padded row of the per-warp scratch (doubles): conflict-free 128-bit reads
#define SF_TILE_DOUBLES (SFGPU_NFIELDS * SF_NT * SF_NT)
#define SF_SCRATCH_DOUBLES (12 * SF_WROW)
#define SF_WARP_SMEM_BYTES ((SF_TILE_DOUBLES + SF_SCRATCH_DOUBLES) * 8)

__device__ __forceinline__ double sf_vacant() { return __longlong_as_double(0x7ff8000000000001LL); }

// sort key of a logical position: tile-major, row-major cells inside the tile
__device__ __forceinline__ unsigned sf_cell_key(const MeshDev &m, double li, double lj, int ntj)
{
    int ci = sf_j2i(li), cj = sf_j2i(lj);
    ci = min(max(ci, 0), m.ni - 2);
    cj = min(max(cj, 0), m.nj - 2);
    const int tile = (ci / SF_TILE) * ntj + (cj / SF_TILE);
    return (unsigned)tile * (SF_TILE * SF_TILE) + (unsigned)((ci % SF_TILE) * SF_TILE + (cj % SF_TILE));
}

struct FastStepArgs {
    const MeshDev *meshes;
    int mesh_id;
    double qm, charge, dt;
    FastPtrs fs;
    const WorkItem *items;
    const unsigned *n_items;
    int ntj;
    RecPtrs exc;              // records list receiving particles that became exceptional
    unsigned long long exc_cap;
    const XferDev *xfer;
    SlowPtrs slow;
    double *dep;
    StepCounters *c;
};

// what happens to a particle of the fast store after sf_move(); shared by the tiled and the tail kernel
__device__ __forceinline__ void fast_epilogue(const FastStepArgs &a, const MeshDev &m, size_t q, int st, bool exact,
                                              const PState &p, const MoveAux &aux, double z0, long long w0bits, int2 tagv_unused,
                                              bool &deposit)
{
    deposit = false;
    if (st == SF_ALIVE && exact && p.dt == 0) {
        a.fs.x[q] = p.x;
        a.fs.y[q] = p.y;
        if (p.z != z0 || p.z != p.z) a.fs.z[q] = p.z;
        a.fs.u[q] = p.u;
        a.fs.v[q] = p.v;
        if (__double_as_longlong(p.w) != w0bits) a.fs.w[q] = p.w;
        deposit = true;
        return;
    }
    // the particle leaves the fast store
    a.fs.mpw[q] = sf_vacant();
    atomicAdd((unsigned long long *)&a.c->fast_delta[a.mesh_id], ~0ULL); // -1
    const int2 tag = a.fs.tag[q];
    if (st == SF_ALIVE) { // stale lc or residual dt: full record, handled by the generic kernel from now on
        const unsigned long long s = atomicAdd(&a.c->n_out[a.mesh_id], 1ULL);
        if (s < a.exc_cap) rec_store(a.exc, s, p, tag);
        else atomicAdd(&a.c->overflow, 1ULL);
        deposit = true;
    } else if (st == SF_DEAD) {
        atomicAdd(&a.c->n_exited, 1ULL);
    } else if (st == SF_REMOVED) {
        atomicAdd(&a.c->n_removed, 1ULL);
    } else if (st == SF_TRANSFER) {
        for (int k = 0; k < 2; k++) {
            if (!(aux.xfer_mask & (1 << k))) continue;
            const int nb = aux.xfer_mesh[k];
            const unsigned long long s = atomicAdd(&a.c->xfer_n[nb], 1ULL);
            if (s < a.xfer[nb].cap) {
                PState cp = p;
                cp.li = aux.xfer_li[k];
                cp.lj = aux.xfer_lj[k];
                rec_store(a.xfer[nb].rec, s, cp, tag);
            } else {
                atomicAdd(&a.c->overflow, 1ULL);
            }
            atomicAdd(&a.c->n_xfer_copies, 1ULL);
        }
    } else { // SF_SLOW
        const unsigned long long s = atomicAdd(&a.c->n_slow, 1ULL);
        if (s < a.slow.cap) {
            rec_store(a.slow.rec, s, p, tag);
            a.slow.old_x[s] = aux.xo; a.slow.old_y[s] = aux.yo;
            a.slow.old_li[s] = aux.lio; a.slow.old_lj[s] = aux.ljo;
            a.slow.bounces[s] = aux.bounces; a.slow.mesh[s] = a.mesh_id;
        } else {
            atomicAdd(&a.c->overflow, 1ULL);
        }
    }
}

// load + move of one fast-store particle.  Returns false for a vacant slot.
__device__ __forceinline__ bool fast_load_move(const FastStepArgs &a, const MeshDev &m, size_t q, PState &p, MoveAux &aux, int &st,
                                               bool &exact, double &z0, long long &w0bits)
{
    p.mpw = a.fs.mpw[q];
    if (p.mpw != p.mpw) return false; // vacant
    p.x = a.fs.x[q]; p.y = a.fs.y[q]; p.z = a.fs.z[q];
    p.u = a.fs.u[q]; p.v = a.fs.v[q]; p.w = a.fs.w[q];
    z0 = p.z;
    w0bits = __double_as_longlong(p.w);
    p.li = (p.x - m.x0) / m.dhx; // UM:158-159: the stored lc of a normal particle is exactly this
    p.lj = (p.y - m.y0) / m.dhy;
    p.dt = 0;
    exact = true;
    const GlobalFieldGather fg;
    st = sf_move(m, a.meshes, a.qm, a.charge, a.dt, false, p, aux, exact, fg);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// tiled kernel: persistent warps pull work items from a queue
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SF_FAST_WARPS * 32)
k_fast_step(FastStepArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *tile = reinterpret_cast<double *>(smem_raw + (size_t)wid * SF_WARP_SMEM_BYTES);
    double *sW = tile + SF_TILE_DOUBLES; // [4][SF_WROW] weights, then [8][SF_WROW] values
    double *sV = sW + 4 * SF_WROW;
    const MeshDev m = a.meshes[a.mesh_id];
    const size_t plane = (size_t)m.ni * m.nj;

    for (int k = lane; k < SF_TILE_DOUBLES; k += 32) tile[k] = 0.0;

    // role of this lane in the reduction: node n (w00,w10,w11,w01) x field f (7 moments + the cell count)
    const int rn = lane >> 3, rf = lane & 7;
    const int noff = (rn == 0) ? 0 : (rn == 1) ? SF_NT : (rn == 2) ? SF_NT + 1 : 1;
    const double *rw = (rf == 7) ? (sV + 7 * SF_WROW) : (sW + rn * SF_WROW); // count lane: 1.0 * 1.0
    const double *rv = sV + rf * SF_WROW;
    double *racc = tile + rf * (SF_NT * SF_NT) + ((rf == 7) ? 0 : noff);
    const bool rflush = (rf < 7) || (rn == 0);

    double sN = 0, sPx = 0, sPy = 0, sPz = 0, sE = 0;
    const unsigned n_items = *a.n_items;
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(&a.c->queue[a.mesh_id], 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const WorkItem wi = a.items[it];
        const int ti0 = (wi.tile / a.ntj) * SF_TILE - SF_HALO; // first node row / column held by the tile
        const int tj0 = (wi.tile % a.ntj) * SF_TILE - SF_HALO;
        for (int b = 0; b < wi.count; b += 32) {
            const bool valid = b + lane < wi.count;
            const size_t q = (size_t)wi.begin + b + lane;
            PState p;
            MoveAux aux;
            int st = SF_REMOVED;
            bool exact = true, present = false, deposit = false;
            double z0 = 0;
            long long w0bits = 0;
            if (valid) present = fast_load_move(a, m, q, p, aux, st, exact, z0, w0bits);
            if (present) fast_epilogue(a, m, q, st, exact, p, aux, z0, w0bits, make_int2(0, 0), deposit);
            // ---- deposit ----
            int key = -1;
            DepW d;
            double val[7];
            if (deposit) {
                sN += p.mpw;
                sPx += p.mpw * p.u;
                sPy += p.mpw * p.v;
                sPz += p.mpw * p.w;
                sE += p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w);
                const bool in = sf_deposit_weights(m, p.li, p.lj, d);
                const int li_ = d.i - ti0, lj_ = d.j - tj0;
                if (in && li_ >= 0 && lj_ >= 0 && li_ < SF_NT - 1 && lj_ < SF_NT - 1) {
                    key = li_ * SF_NT + lj_;
                    sf_deposit_values(p, val);
                } else {
                    deposit_global(m, p, a.dep);
                    atomicAdd(&a.c->n_fallback, 1ULL);
                }
            }
            if (key >= 0) {
                sW[0 * SF_WROW + lane] = d.w00;
                sW[1 * SF_WROW + lane] = d.w10;
                sW[2 * SF_WROW + lane] = d.w11;
                sW[3 * SF_WROW + lane] = d.w01;
#pragma unroll
                for (int f = 0; f < 7; f++) sV[f * SF_WROW + lane] = val[f];
                sV[7 * SF_WROW + lane] = 1.0;
            } else {
#pragma unroll
                for (int f = 0; f < 4; f++) sW[f * SF_WROW + lane] = 0.0;
#pragma unroll
                for (int f = 0; f < 8; f++) sV[f * SF_WROW + lane] = 0.0;
            }
            const int prev = __shfl_up_sync(0xffffffffu, key, 1);
            const unsigned bmask = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
            const unsigned any = __ballot_sync(0xffffffffu, key >= 0);
            __syncwarp();
            if (any) {
                double acc = 0.0;
                int cur = __shfl_sync(0xffffffffu, key, 0);
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const double2 w2 = *reinterpret_cast<const double2 *>(rw + k);
                    const double2 v2 = *reinterpret_cast<const double2 *>(rv + k);
                    if (k > 0 && ((bmask >> k) & 1u)) {
                        if (cur >= 0 && rflush) racc[cur] += acc;
                        acc = 0.0;
                        cur = __shfl_sync(0xffffffffu, key, k);
                    }
                    acc = __fma_rn(w2.x, v2.x, acc);
                    if ((bmask >> (k + 1)) & 1u) {
                        if (cur >= 0 && rflush) racc[cur] += acc;
                        acc = 0.0;
                        cur = __shfl_sync(0xffffffffu, key, k + 1);
                    }
                    acc = __fma_rn(w2.y, v2.y, acc);
                }
                if (cur >= 0 && rflush) racc[cur] += acc;
            }
            __syncwarp();
        }
        // ---- add the warp tile to the global deposit and clear it ----
        for (int k = lane; k < SF_TILE_DOUBLES; k += 32) {
            const double v = tile[k];
            if (v != 0.0) {
                const int f = k / (SF_NT * SF_NT), r = k % (SF_NT * SF_NT);
                const int gi = ti0 + r / SF_NT, gj = tj0 + r % SF_NT;
                if (gi >= 0 && gj >= 0 && gi < m.ni && gj < m.nj) atomicAdd(a.dep + f * plane + (size_t)gi * m.nj + gj, v);
                tile[k] = 0.0;
            }
        }
        __syncwarp();
    }
    sN = warp_sum(sN); sPx = warp_sum(sPx); sPy = warp_sum(sPy); sPz = warp_sum(sPz); sE = warp_sum(sE);
    if (lane == 0 && sN != 0) {
        atomicAdd(&a.c->sums[0], sN); atomicAdd(&a.c->sums[1], sPx); atomicAdd(&a.c->sums[2], sPy);
        atomicAdd(&a.c->sums[3], sPz); atomicAdd(&a.c->sums[4], sE);
    }
}

// ---------------------------------------------------------------------------------------------------------
// tail kernel: fast-store particles appended since the last sort (injection) -- same arithmetic, global deposit
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fast_tail(FastStepArgs a, unsigned long long first, unsigned long long n)
{
    const MeshDev m = a.meshes[a.mesh_id];
    double sN = 0, sPx = 0, sPy = 0, sPz = 0, sE = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n; q0 += stride) {
        const size_t q = first + q0;
        PState p;
        MoveAux aux;
        int st = SF_REMOVED;
        bool exact = true, deposit = false;
        double z0 = 0;
        long long w0bits = 0;
        if (!fast_load_move(a, m, q, p, aux, st, exact, z0, w0bits)) continue;
        fast_epilogue(a, m, q, st, exact, p, aux, z0, w0bits, make_int2(0, 0), deposit);
        if (deposit) {
            deposit_global(m, p, a.dep);
            sN += p.mpw;
            sPx += p.mpw * p.u;
            sPy += p.mpw * p.v;
            sPz += p.mpw * p.w;
            sE += p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w);
        }
    }
    sN = warp_sum(sN); sPx = warp_sum(sPx); sPy = warp_sum(sPy); sPz = warp_sum(sPz); sE = warp_sum(sE);
    if ((threadIdx.x & 31) == 0 && sN != 0) {
        atomicAdd(&a.c->sums[0], sN); atomicAdd(&a.c->sums[1], sPx); atomicAdd(&a.c->sums[2], sPy);
        atomicAdd(&a.c->sums[3], sPz); atomicAdd(&a.c->sums[4], sE);
    }
}

// ---------------------------------------------------------------------------------------------------------
// cell sort + compaction (K3): counting sort by cell key, out of place
// ---------------------------------------------------------------------------------------------------------
#define SF_KEY_NONE 0xffffffffu

// pass 1: key and rank of every particle; hist[key] = particles per cell.  Vacant slots get SF_KEY_NONE.
__global__ void __launch_bounds__(256)
k_sort_count(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs fs, unsigned long long n, int ntj, unsigned *__restrict__ hist,
             unsigned *__restrict__ keys, unsigned *__restrict__ ranks)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    unsigned key = SF_KEY_NONE;
    if (q < n) {
        const double mpw = fs.mpw[q];
        if (mpw == mpw) key = sf_cell_key(m, (fs.x[q] - m.x0) / m.dhx, (fs.y[q] - m.y0) / m.dhy, ntj);
    }
    const unsigned act = __ballot_sync(0xffffffffu, key != SF_KEY_NONE);
    if (key != SF_KEY_NONE) {
        const unsigned grp = __match_any_sync(act, key);
        const int leader = __ffs(grp) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&hist[key], (unsigned)__popc(grp));
        base = __shfl_sync(grp, base, leader);
        ranks[q] = base + __popc(grp & ((1u << lane) - 1u));
    }
    if (q < n) keys[q] = key;
}

// pass 3: scatter to the sorted position
__global__ void __launch_bounds__(256)
k_sort_scatter(FastPtrs in, FastPtrs out, unsigned long long n, const unsigned *__restrict__ offs, const unsigned *__restrict__ keys,
               const unsigned *__restrict__ ranks)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const unsigned key = keys[q];
    if (key == SF_KEY_NONE) return;
    const size_t d = (size_t)offs[key] + ranks[q];
    out.x[d] = in.x[q]; out.y[d] = in.y[q]; out.z[d] = in.z[q];
    out.u[d] = in.u[q]; out.v[d] = in.v[q]; out.w[d] = in.w[q];
    out.mpw[d] = in.mpw[q];
    out.tag[d] = in.tag[q];
}

// work items: each tile's run [offs[tile*64], offs[(tile+1)*64]) cut into pieces of <= SF_ITEM_MAX particles
__global__ void k_build_items(const unsigned *__restrict__ offs, int n_tiles, WorkItem *__restrict__ items, unsigned *__restrict__ n_items,
                              unsigned max_items)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const unsigned b = offs[(size_t)t * SF_TILE * SF_TILE], e = offs[(size_t)(t + 1) * SF_TILE * SF_TILE];
    if (e <= b) return;
    const unsigned cnt = e - b, pieces = (cnt + SF_ITEM_MAX - 1) / SF_ITEM_MAX;
    const unsigned per = ((cnt + pieces - 1) / pieces + 31u) & ~31u; // whole warps, balanced
    const unsigned s = atomicAdd(n_items, pieces);
    for (unsigned k = 0; k < pieces && s + k < max_items; k++) {
        const unsigned pb = b + k * per, pe = min(e, pb + per);
        WorkItem w;
        w.begin = pb;
        w.count = pe > pb ? (int)(pe - pb) : 0;
        w.tile = t;
        items[s + k] = w;
    }
}

// injection into the fast store (KM:759-802 for the common case lc == null): lc = XtoL(pos), -0.5dt rewind, dt = 0.
// Particles the fast store cannot represent (plus-edge clamp KM:770-773, NaN weight) go to the record list.
__global__ void __launch_bounds__(256)
k_inject_fast(const MeshDev *__restrict__ meshes, int mesh_id, double qm, double dt_step, int rewind, FastPtrs fs, unsigned long long first,
              unsigned long long n, RecPtrs rec, unsigned long long rec_first, unsigned long long rec_cap, StepCounters *__restrict__ c)
{
    const MeshDev m = meshes[mesh_id];
    const GlobalFieldGather fg;
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n) return;
    const size_t q = first + q0;
    PState p;
    p.x = fs.x[q]; p.y = fs.y[q]; p.z = fs.z[q]; p.u = fs.u[q]; p.v = fs.v[q]; p.w = fs.w[q]; p.mpw = fs.mpw[q];
    p.li = (p.x - m.x0) / m.dhx;
    p.lj = (p.y - m.y0) / m.dhy;
    p.dt = 0;
    bool normal = p.mpw == p.mpw;
    if (p.li >= m.ni) { p.li = m.ni - 1; normal = false; }
    if (p.lj >= m.nj) { p.lj = m.nj - 1; normal = false; }
    if (rewind) sf_kick(m, qm, -0.5 * dt_step, p, fg);
    if (!(isfinite(p.u) && isfinite(p.v) && isfinite(p.w))) { // MeshData.addParticle drops it, KM:1357-1361
        atomicAdd(&c->n_bad, 1ULL);
        fs.mpw[q] = sf_vacant();
        return;
    }
    if (normal) {
        fs.u[q] = p.u; fs.v[q] = p.v; fs.w[q] = p.w;
        return;
    }
    const unsigned long long s = rec_first + atomicAdd(&c->n_exc[mesh_id], 1ULL);
    if (s < rec_cap) rec_store(rec, s, p, fs.tag[q]);
    else atomicAdd(&c->overflow, 1ULL);
    fs.mpw[q] = sf_vacant();
}

// fast store -> full records (download / restart / iterators): lc = XtoL(pos), dt = 0
__global__ void __launch_bounds__(256)
k_fast_to_records(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs fs, unsigned long long first, unsigned long long n, RecPtrs out)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n) return;
    const size_t q = first + q0;
    PState p;
    p.x = fs.x[q]; p.y = fs.y[q]; p.z = fs.z[q]; p.u = fs.u[q]; p.v = fs.v[q]; p.w = fs.w[q]; p.mpw = fs.mpw[q];
    p.li = (p.x - m.x0) / m.dhx;
    p.lj = (p.y - m.y0) / m.dhy;
    p.dt = 0;
    rec_store(out, q0, p, fs.tag[q]);
}

// full records -> fast store slots (upload after a host-side mutation); a record the fast store cannot hold
// exactly (lc != XtoL(pos) or dt != 0) is appended to the record list instead and its slot vacated
__global__ void __launch_bounds__(256)
k_records_to_fast(const MeshDev *__restrict__ meshes, int mesh_id, RecPtrs in, unsigned long long n, FastPtrs fs, unsigned long long first,
                  RecPtrs rec, unsigned long long rec_first, unsigned long long rec_cap, StepCounters *__restrict__ c)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n) return;
    const size_t q = first + q0;
    PState p;
    int2 tag;
    rec_load(in, q0, p, tag);
    const bool normal = p.mpw == p.mpw && p.dt == 0 && p.li == (p.x - m.x0) / m.dhx && p.lj == (p.y - m.y0) / m.dhy;
    fs.tag[q] = tag;
    if (normal) {
        fs.x[q] = p.x; fs.y[q] = p.y; fs.z[q] = p.z; fs.u[q] = p.u; fs.v[q] = p.v; fs.w[q] = p.w; fs.mpw[q] = p.mpw;
        return;
    }
    const unsigned long long s = rec_first + atomicAdd(&c->n_exc[mesh_id], 1ULL);
    if (s < rec_cap) rec_store(rec, s, p, tag);
    else atomicAdd(&c->overflow, 1ULL);
    fs.mpw[q] = sf_vacant();
}
