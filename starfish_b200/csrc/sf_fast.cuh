// sf_fast.cuh -- the B200 fast path: fused gather / push / locate / deposit over the cell-sorted fast store.
//
// Layout (DESIGN.md "Data layout"): structure of arrays x,y,z,u,v,w,mpw (+tag), sorted by cell inside SF_TILE x SF_TILE-cell (4 x 4)
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

#ifndef SF_FAST_WARPS
#define SF_FAST_WARPS 5 // warps per CTA of the tiled kernel, two-cell halo (4 CTAs of 5 warps x 10.5 KB of shared memory: 20 warps / SM, 96 registers)
#endif
#ifndef SF_PPT
#define SF_PPT 1        // particles per lane and batch (independent instruction streams hide FP64 latency)
#endif
#ifndef SF_FAST_WARPS_H1
#define SF_FAST_WARPS_H1 5 // ... and of its one-cell-halo instance (5 warps x 8.5 KB; 20 warps / SM is the register limit either way)
#endif
#ifndef SF_FAST_MIN_CTAS
#define SF_FAST_MIN_CTAS 4
#endif
#define SF_WROW (32 * SF_PPT + 2) // padded row of the per-warp scratch (doubles): conflict-free 128-bit reads
#define SF_TILE_DOUBLES (SFGPU_NFIELDS * SF_NT * SF_NT)
#ifndef SF_STAGE
#define SF_STAGE 0 // 1: cp.async prefetch of the next batch through shared memory; 0: plain loads, 3.5 KB less per warp
#endif
#ifndef SF_ETILE
#define SF_ETILE 1 // 1: the E field of the work item's tile (+ SF_EHALO cells) is staged in shared memory, the gathers of the common path read it
#endif
#ifndef SF_EHALO
#define SF_EHALO 2
#endif
#define SF_ENT (SF_TILE + 2 * SF_EHALO + 1) // nodes per edge of the staged E tile
#define SF_ETILE_DOUBLES (SF_ETILE * (2 * SF_ENT * SF_ENT + (2 * SF_ENT * SF_ENT) % 2))
#define SF_EXTRA 6 // per-warp sums next to the tile: energy, fallback N/Px/Py/Pz/E (the fallback count rides in the tag of E)
#define SF_SCRATCH_DOUBLES (SF_EXTRA + 13 * SF_WROW + 16 + 32 * SF_PPT + 16 + 16 * SF_PPT + 2 + SF_ETILE_DOUBLES + SF_STAGE * 2 * 7 * 32 * SF_PPT)
#define SF_WARP_SMEM_BYTES ((SF_TILE_DOUBLES + SF_SCRATCH_DOUBLES) * 8)
// geometry of the tiled kernel per halo width.  A halo of one cell is enough while the deposits of a particle stay within one cell of the
// cell it was sorted into (the sort key is the PREDICTED cell, so that is one step of thermal spread either way): the smaller tile lets
// 15 warps share an SM instead of 12.  Populations that move further per step fall back to the two-cell halo (sfgpu_step picks, SFGPU_HALO)
template <int HALO> struct FastGeom {
    static constexpr int NT = SF_TILE + 2 * HALO + 1;              // nodes per edge of the accumulation tile
    static constexpr int TILE_DOUBLES = SFGPU_NFIELDS * NT * NT;
    static constexpr int WARP_BYTES = (TILE_DOUBLES + SF_SCRATCH_DOUBLES) * 8;
    static constexpr int WARPS = (HALO == 1) ? SF_FAST_WARPS_H1 : SF_FAST_WARPS;   // warps per CTA; SF_FAST_MIN_CTAS CTAs per SM
    static_assert(WARP_BYTES % 16 == 0 && (TILE_DOUBLES + SF_EXTRA) % 2 == 0, "128-bit shared-memory operands");
    static_assert(SF_FAST_MIN_CTAS * WARPS * WARP_BYTES <= 227 * 1024, "shared memory of an SM");
};

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
    MeshDev m;                // the mesh of this launch, by value: kernel parameters live in the constant bank
    const MeshDev *meshes;    // all meshes (neighbour lookup of the MESH hand-off)
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
    unsigned *defer;          // slots k_fast_step leaves to k_fast_deferred: bit 31 clear = re-run the particle through sf_move() from its
    unsigned defer_cap;       // stored state; bit 31 set = already pushed and stored, only its deposit (it left the warp tile) is due
    // the step before a cell sort does the sort's counting pass on the way (k_fast_step<.., PREP = true>): key of the cell each particle is predicted to
    // occupy sort_dt from now, its rank inside that cell, and the histogram -- k_sort_count's outputs without its pass over the particles
    unsigned *sort_keys, *sort_ranks, *sort_hist;
    double sort_dt;
};
#define SF_DEFER_DEPOSIT_ONLY 0x80000000u
#define SF_KEY_NONE 0xffffffffu // sort key of a vacant slot

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
    } else if (st == SF_ABSORBED) {
        atomicAdd(&a.c->n_absorbed, 1ULL);
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
// straight-line common case of sf_move(): one substep, no B field, no segments, the particle starts and ends
// strictly inside the mesh.  No branches, so the compiler interleaves the SF_PPT independent particles a lane
// carries (FP64 latency hiding).  Every expression is the one sf_move() evaluates, in the same order, so the two
// paths are bit-identical; `ok` false means "not the common case": the caller re-runs the particle through
// sf_move() from its original state.
// ---------------------------------------------------------------------------------------------------------
// where the common path finds the four E nodes of a cell: global memory (L1 / L2), or the tile staged in shared memory
struct EGlobal {
    __device__ __forceinline__ bool corners(const MeshDev &m, int i, int j, bool ok, double (&e)[8]) const
    {
        const int ic = min(max(i, 0), m.ni - 2), jc = min(max(j, 0), m.nj - 2); // safe addresses when !ok
        const size_t n00 = (size_t)ic * m.nj + jc;
        const double *fi = m.efi + n00, *fj = m.efj + n00;
        e[0] = __ldg(fi); e[1] = __ldg(fi + m.nj); e[2] = __ldg(fi + m.nj + 1); e[3] = __ldg(fi + 1);
        e[4] = __ldg(fj); e[5] = __ldg(fj + m.nj); e[6] = __ldg(fj + m.nj + 1); e[7] = __ldg(fj + 1);
        return ok;
    }
};
struct ETile { // efi, efj over SF_ENT x SF_ENT nodes starting at node (i0, j0); a particle outside the tile is not the common case
    const double *e;
    int i0, j0;
    __device__ __forceinline__ bool corners(const MeshDev &m, int i, int j, bool ok, double (&c)[8]) const
    {
        const int ti = i - i0, tj = j - j0;
        ok = ok && (unsigned)ti < (unsigned)(SF_ENT - 1) && (unsigned)tj < (unsigned)(SF_ENT - 1);
        const double *b = e + (ok ? ti * SF_ENT + tj : 0);
        c[0] = b[0]; c[1] = b[SF_ENT]; c[2] = b[SF_ENT + 1]; c[3] = b[1];
        c[4] = b[SF_ENT * SF_ENT]; c[5] = b[SF_ENT * SF_ENT + SF_ENT]; c[6] = b[SF_ENT * SF_ENT + SF_ENT + 1]; c[7] = b[SF_ENT * SF_ENT + 1];
        return ok;
    }
};

template <bool SEG, class EField> // SEG: the mesh has segment nodes; a sub-step whose node box touches one is not the common case
__device__ __forceinline__ bool sf_move_simple(const MeshDev &m, double qm, double dt, PState &p, const EField &ef)
{
    const int ni = m.ni, nj = m.nj;
    const int i = sf_j2i(p.li), j = sf_j2i(p.lj);
    bool ok = (p.mpw > 0) && i >= 0 && j >= 0 && i < ni - 1 && j < nj - 1;
    const double di = p.li - i, dj = p.lj - j;
    double e[8];
    ok = ef.corners(m, i, j, ok, e);
    const double w00 = (1 - di) * (1 - dj), w10 = di * (1 - dj), w11 = di * dj, w01 = (1 - di) * dj; // F2D:345-348
    double ex = w00 * e[0];
    ex += w10 * e[1];
    ex += w11 * e[2];
    ex += w01 * e[3];
    double ey = w00 * e[4];
    ey += w10 * e[5];
    ey += w11 * e[6];
    ey += w01 * e[7];
    const double u = p.u + qm * ex * dt; // KM:345-346 with part.dt = 0 + dt
    const double v = p.v + qm * ey * dt;
    double x = p.x + u * dt; // KM:369-370
    double y = p.y + v * dt;
    double z, un = u, vn = v, wn = p.w;
    if (m.domain == SFGPU_XY) {
        z = p.z + p.w * dt; // KM:380
    } else if (m.domain == SFGPU_RZ) { // KM:424-442
        const double A = p.w * dt, B = x, R = sqrt(A * A + B * B);
        const double c = B / R, s = A / R;
        z = p.z - asin(s);
        x = R;
        un = c * u + s * p.w;
        wn = -s * u + c * p.w;
    } else { // KM:444-462
        const double A = p.w * dt, B = y, R = sqrt(A * A + B * B);
        const double c = B / R, s = A / R;
        z = p.z + acos(c);
        y = R;
        vn = c * v + s * p.w;
        wn = -s * v + c * p.w;
    }
    const double li = sf_div_exact(x - m.x0, m.dhx, m.rdhx, m.fastdiv); // UM:158-159
    const double lj = sf_div_exact(y - m.y0, m.dhy, m.rdhy, m.fastdiv);
    ok = ok && li >= 0 && lj >= 0 && li < m.nim1 && lj < m.njm1; // KM:606 (a NaN takes the general path)
    if (SEG && ok) { // KM:482-518: a segment node in the node box of the sub-step sends it through ProcessBoundary proper
        const int i2 = sf_j2i(li), j2 = sf_j2i(lj);
        const uint8_t *hs = m.has_seg;
        ok = abs(i2 - i) <= 1 && abs(j2 - j) <= 1 &&
             !(hs[(size_t)i * nj + j] | hs[(size_t)i2 * nj + j] | hs[(size_t)i * nj + j2] | hs[(size_t)i2 * nj + j2]);
    }
    if (ok) {
        p.x = x; p.y = y; p.z = z; p.u = un; p.v = vn; p.w = wn; p.li = li; p.lj = lj;
    }
    return ok;
}

// asynchronous global -> shared copy of one batch (7 state doubles per particle, each lane fetches the slots it will
// read back itself, so no cross-lane synchronisation is needed): LDGSTS, no registers held while in flight
__device__ __forceinline__ void sf_prefetch_batch(const FastPtrs &fs, double *stage, size_t begin, int b, int count, int lane)
{
    const double *src[7] = {fs.x, fs.y, fs.z, fs.u, fs.v, fs.w, fs.mpw};
#pragma unroll
    for (int j = 0; j < SF_PPT; j++) {
        const int o = b + j * 32 + lane;
        if (o < count) {
#pragma unroll
            for (int f = 0; f < 7; f++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + f * 32 * SF_PPT + j * 32 + lane);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src[f] + begin + o) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// tiled kernel: persistent warps pull work items from a queue; every lane carries SF_PPT particles per batch
// ---------------------------------------------------------------------------------------------------------
template <bool SEG, int HALO, bool PREP> // PREP: also writes the counting pass of the cell sort that follows this step (FastStepArgs::sort_*); HALO: cells kept around the SF_TILE x SF_TILE tile in the warp-private accumulation tile
__global__ void __launch_bounds__(FastGeom<HALO>::WARPS * 32, SF_FAST_MIN_CTAS)
k_fast_step(const __grid_constant__ FastStepArgs a, const FastStepArgs *__restrict__ ga)
{
    constexpr int NT = FastGeom<HALO>::NT, TD = FastGeom<HALO>::TILE_DOUBLES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *tile = reinterpret_cast<double *>(smem_raw + (size_t)wid * FastGeom<HALO>::WARP_BYTES);
    double *sW = tile + TD + SF_EXTRA; // tile, extra sums of the warp (energy, fallback N/P/E), then [4][SF_WROW] weights, [9][SF_WROW] values
    double *sV = sW + 4 * SF_WROW;
    // weight and value of the counting lanes (f == 7): a row of ones placed on the bank group of value row 7 (rows are 272 bytes apart, i.e. 4 banks
    // further per row: the eight lanes of a quarter warp then read eight different bank groups and the 128-bit operand loads stay conflict free)
    double *sOnes = sV + 8 * SF_WROW + ((7 * SF_WROW - 8 * SF_WROW) % 16 + 16) % 16;
    // ... and a second row of ones for the WEIGHT operand of the counting lanes, on the bank group of value row 0 = weight row 4: none of the
    // four weight rows a quarter warp reads (the first one collided with weight row 3)
    double *sOnesW = sV + 9 * SF_WROW + 16 + ((0 * SF_WROW - (9 * SF_WROW + 16)) % 16 + 16) % 16;
    int *sKey = reinterpret_cast<int *>(sV + 9 * SF_WROW + 16 + 32 * SF_PPT + 16); // [32 * SF_PPT] tile-local cell of each row, then the fallback count
#if SF_ETILE
    double *sE = reinterpret_cast<double *>(sKey + 32 * SF_PPT + 4); // [2][SF_ENT][SF_ENT] efi, efj around the tile
#endif
#if SF_STAGE
    double *sIn = reinterpret_cast<double *>(sKey + 32 * SF_PPT + 4) + SF_ETILE_DOUBLES; // [2 stages][7][32 * SF_PPT] prefetched particle state
#endif
    const MeshDev &m = a.m;
    const size_t plane = (size_t)m.ni * m.nj;
    const bool simple_ok = !m.has_b && a.dt > 0 && (SEG || !m.any_seg);

    for (int k = lane; k < TD + SF_EXTRA; k += 32) tile[k] = 0.0;
    if (lane == 0) sKey[32 * SF_PPT] = 0;
    for (int k = lane; k < 32 * SF_PPT; k += 32) sOnes[k] = sOnesW[k] = 1.0;
    __syncwarp();

    // role of this lane in the reduction: node n (w00,w10,w11,w01) x field f (7 moments; f == 7: n == 0 counts the
    // particles of the cell (mpc), n == 1 sums mpw*|vel| of the whole work item (energy sum, KM:412))
    const int rn = lane >> 3, rf = lane & 7;
    const int noff = (rn == 0) ? 0 : (rn == 1) ? NT : (rn == 2) ? NT + 1 : 1;
    const double *rw = (rf == 7) ? sOnesW : (sW + rn * SF_WROW); // f == 7 lanes: weight 1.0
    const double *rv = (rf == 7 && rn != 1) ? sOnes : (sV + rf * SF_WROW); // row 7 of sV = mpw*|vel|
    const bool renergy = rf == 7 && rn == 1;
    double *racc = renergy ? (tile + TD) : (tile + rf * (NT * NT) + ((rf == 7) ? 0 : noff));
    const int rmul = renergy ? 0 : 1;
    const bool rflush = (rf < 7) || (rn <= 1);

    const unsigned n_items = *a.n_items;
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(&a.c->queue[a.mesh_id], 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const WorkItem wi = a.items[it];
        const int ti0 = (wi.tile / a.ntj) * SF_TILE - HALO; // first node row / column held by the tile
        const int tj0 = (wi.tile % a.ntj) * SF_TILE - HALO;
#if SF_ETILE
        // E field of the tile's neighbourhood (F2D:300-350 reads four nodes per field and particle): staged once per work item
        ETile etile;
        etile.e = sE;
        etile.i0 = ti0 + HALO - SF_EHALO;
        etile.j0 = tj0 + HALO - SF_EHALO;
        __syncwarp();
        for (int k = lane; k < SF_ENT * SF_ENT; k += 32) {
            const int gi = etile.i0 + k / SF_ENT, gj = etile.j0 + k % SF_ENT;
            double fi = 0.0, fj = 0.0;
            if (gi >= 0 && gj >= 0 && gi < m.ni && gj < m.nj) {
                fi = __ldg(m.efi + (size_t)gi * m.nj + gj);
                fj = __ldg(m.efj + (size_t)gi * m.nj + gj);
            }
            sE[k] = fi;
            sE[SF_ENT * SF_ENT + k] = fj;
        }
        __syncwarp();
#endif
#if SF_STAGE
        sf_prefetch_batch(a.fs, sIn, (size_t)wi.begin, 0, wi.count, lane);
#else
        // the state of the NEXT batch is fetched into registers while the current one is processed (the kernel has registers to spare
        // since the rare paths left it): the first use of a batch no longer waits for HBM
        double nx[SF_PPT][7];
#pragma unroll
        for (int j = 0; j < SF_PPT; j++) {
            const int o = j * 32 + lane;
            nx[j][6] = sf_vacant();
            if (o < wi.count) {
                const size_t q = (size_t)wi.begin + o;
                nx[j][0] = a.fs.x[q]; nx[j][1] = a.fs.y[q]; nx[j][2] = a.fs.z[q];
                nx[j][3] = a.fs.u[q]; nx[j][4] = a.fs.v[q]; nx[j][5] = a.fs.w[q];
                nx[j][6] = a.fs.mpw[q];
            }
        }
#endif
        for (int b = 0; b < wi.count; b += 32 * SF_PPT) {
            PState p[SF_PPT];
            bool present[SF_PPT], done[SF_PPT];
            double z0[SF_PPT];
            long long w0bits[SF_PPT];
            // ---- the batch was prefetched into shared memory with cp.async while the previous one was processed;
            //      start the next one now (global-memory latency hidden behind ~700 instructions of work) ----
#if SF_STAGE
            const int stage = (b / (32 * SF_PPT)) & 1;
            const bool more = b + 32 * SF_PPT < wi.count;
            if (more) sf_prefetch_batch(a.fs, sIn + (stage ^ 1) * (7 * 32 * SF_PPT), (size_t)wi.begin, b + 32 * SF_PPT, wi.count, lane);
            if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
#pragma unroll
            for (int j = 0; j < SF_PPT; j++) {
                const int o = b + j * 32 + lane;
                present[j] = o < wi.count;
                p[j].mpw = sf_vacant();
                if (present[j]) {
#if SF_STAGE
                    const double *in = sIn + stage * (7 * 32 * SF_PPT) + j * 32 + lane;
                    p[j].x = in[0 * 32 * SF_PPT]; p[j].y = in[1 * 32 * SF_PPT]; p[j].z = in[2 * 32 * SF_PPT];
                    p[j].u = in[3 * 32 * SF_PPT]; p[j].v = in[4 * 32 * SF_PPT]; p[j].w = in[5 * 32 * SF_PPT];
                    p[j].mpw = in[6 * 32 * SF_PPT];
#else
                    p[j].x = nx[j][0]; p[j].y = nx[j][1]; p[j].z = nx[j][2];
                    p[j].u = nx[j][3]; p[j].v = nx[j][4]; p[j].w = nx[j][5];
                    p[j].mpw = nx[j][6];
#endif
                }
#if !SF_STAGE
                nx[j][6] = sf_vacant();
                if (o + 32 * SF_PPT < wi.count) {
                    const size_t q = (size_t)wi.begin + o + 32 * SF_PPT;
                    nx[j][0] = a.fs.x[q]; nx[j][1] = a.fs.y[q]; nx[j][2] = a.fs.z[q];
                    nx[j][3] = a.fs.u[q]; nx[j][4] = a.fs.v[q]; nx[j][5] = a.fs.w[q];
                    nx[j][6] = a.fs.mpw[q];
                }
#endif
            }
            // ---- common case, branch free ----
#pragma unroll
            for (int j = 0; j < SF_PPT; j++) {
                present[j] = present[j] && (p[j].mpw == p[j].mpw);
                z0[j] = p[j].z;
                w0bits[j] = __double_as_longlong(p[j].w);
                p[j].li = sf_div_exact(p[j].x - m.x0, m.dhx, m.rdhx, m.fastdiv); // the stored lc of a normal particle is exactly XtoL(pos)
                p[j].lj = sf_div_exact(p[j].y - m.y0, m.dhy, m.rdhy, m.fastdiv);
                p[j].dt = 0;
#if SF_ETILE
                done[j] = present[j] && simple_ok && sf_move_simple<SEG>(m, a.qm, a.dt, p[j], etile);
#else
                done[j] = present[j] && simple_ok && sf_move_simple<SEG>(m, a.qm, a.dt, p[j], EGlobal());
#endif
            }
            int key[SF_PPT];
            DepW dw[SF_PPT];
#pragma unroll
            for (int j = 0; j < SF_PPT; j++) {
                const size_t q = (size_t)wi.begin + b + j * 32 + lane;
                bool deposit = false;
                if (done[j]) { // alive, exact lc, dt == 0: in-place store
                    a.fs.x[q] = p[j].x;
                    a.fs.y[q] = p[j].y;
                    if (p[j].z != z0[j]) a.fs.z[q] = p[j].z;
                    a.fs.u[q] = p[j].u;
                    a.fs.v[q] = p[j].v;
                    if (__double_as_longlong(p[j].w) != w0bits[j]) a.fs.w[q] = p[j].w;
                    deposit = true;
                } else if (present[j]) { // anything else (boundaries, segments, removal ...): k_fast_deferred re-runs it from the stored state
                    const unsigned long long s_ = atomicAdd(&a.c->n_defer[a.mesh_id], 1ULL);
                    if (s_ < a.defer_cap) a.defer[s_] = (unsigned)q;
                }
                // ---- deposit: weights and the tile-local cell of this particle ----
                key[j] = -1;
                if (deposit) {
                    const bool in = sf_deposit_weights(m, p[j].li, p[j].lj, dw[j]);
                    const int li_ = dw[j].i - ti0, lj_ = dw[j].j - tj0;
                    if (in && li_ >= 0 && lj_ >= 0 && li_ < NT - 1 && lj_ < NT - 1) {
                        key[j] = li_ * NT + lj_;
                    } else { // pushed and stored, but its deposit misses the warp tile: left to k_fast_deferred
                        const unsigned long long s_ = atomicAdd(&a.c->n_defer[a.mesh_id], 1ULL);
                        if (s_ < a.defer_cap) a.defer[s_] = (unsigned)q | SF_DEFER_DEPOSIT_ONLY;
                    }
                }
            }
            if (PREP) { // counting pass of the next sort: same key arithmetic as k_sort_count, on the state this step just stored
#pragma unroll
                for (int j = 0; j < SF_PPT; j++) {
                    const int o = b + j * 32 + lane;
                    const size_t q = (size_t)wi.begin + o;
                    unsigned skey = SF_KEY_NONE;
                    if (done[j]) {
                        double xs = p[j].x, ys = p[j].y;
                        if (a.sort_dt != 0) { xs += p[j].u * a.sort_dt; ys += p[j].v * a.sort_dt; }
                        skey = sf_cell_key(m, sf_div_exact(xs - m.x0, m.dhx, m.rdhx, m.fastdiv), sf_div_exact(ys - m.y0, m.dhy, m.rdhy, m.fastdiv), a.ntj);
                    }
                    const unsigned act = __ballot_sync(0xffffffffu, skey != SF_KEY_NONE);
                    if (skey != SF_KEY_NONE) {
                        const unsigned sg = __match_any_sync(act, skey);
                        const int sl = __ffs(sg) - 1;
                        unsigned sbase = 0;
                        if (lane == sl) sbase = atomicAdd(&a.sort_hist[skey], (unsigned)__popc(sg));
                        sbase = __shfl_sync(sg, sbase, sl);
                        a.sort_ranks[q] = sbase + __popc(sg & ((1u << lane) - 1u));
                    }
                    if (o < wi.count) a.sort_keys[q] = skey; // vacant slots and deferred particles: none (k_sort_count_fix keys the deferred ones once they are finished)
                }
            }
            // ---- group the 32 particles of each set by cell inside the warp: row = position in cell order ----
            unsigned bmask[SF_PPT];
#pragma unroll
            for (int j = 0; j < SF_PPT; j++) {
                const unsigned grp = __match_any_sync(0xffffffffu, key[j]);
                const int leader = __ffs(grp) - 1;
                const int rank = __popc(grp & ((1u << lane) - 1u));
                const bool isl = lane == leader;
                // first row of this lane's group = number of particles whose group leader is a lower lane: five ballots over the bits of the
                // leader (a radix rank) instead of a five-step shuffle scan over the group sizes: no dependent shuffle chain
                int off = 0;
                {
                    unsigned eq = 0xffffffffu; // lanes whose leader agrees with mine on the bits seen so far
#pragma unroll
                    for (int b = 4; b >= 0; b--) {
                        const unsigned bal = __ballot_sync(0xffffffffu, (leader >> b) & 1);
                        if ((leader >> b) & 1) {
                            off += __popc(eq & ~bal);
                            eq &= bal;
                        } else {
                            eq &= ~bal;
                        }
                    }
                }
                const int row = j * 32 + off + rank;
                bmask[j] = __reduce_or_sync(0xffffffffu, isl ? (1u << off) : 0u);
                sKey[row] = key[j];
                if (key[j] >= 0) {
                    double val[7];
                    sf_deposit_values(p[j], val);
                    sW[0 * SF_WROW + row] = dw[j].w00;
                    sW[1 * SF_WROW + row] = dw[j].w10;
                    sW[2 * SF_WROW + row] = dw[j].w11;
                    sW[3 * SF_WROW + row] = dw[j].w01;
#pragma unroll
                    for (int f = 0; f < 7; f++) sV[f * SF_WROW + row] = val[f];
                    sV[7 * SF_WROW + row] = p[j].mpw * sqrt(p[j].u * p[j].u + p[j].v * p[j].v + p[j].w * p[j].w); // KM:412
                } else {
#pragma unroll
                    for (int f = 0; f < 4; f++) sW[f * SF_WROW + row] = 0.0;
#pragma unroll
                    for (int f = 0; f < 8; f++) sV[f * SF_WROW + row] = 0.0;
                }
            }
            __syncwarp();
            // ---- lane (n,f) walks the rows; the running sum is added to the private tile when the cell changes ----
            double acc = 0.0;
            int cur = -1;
#pragma unroll
            for (int j = 0; j < SF_PPT; j++) {
#pragma unroll
                for (int k0 = 0; k0 < 32; k0 += 8) {
                    double2 w2[4], v2[4]; // operands of 8 rows fetched up front: one shared-memory latency per chunk
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        w2[h] = *reinterpret_cast<const double2 *>(rw + j * 32 + k0 + 2 * h);
                        v2[h] = *reinterpret_cast<const double2 *>(rv + j * 32 + k0 + 2 * h);
                    }
#pragma unroll
                    for (int g = 0; g < 2; g++) { // 4 rows per test: most groups of 4 hold no cell boundary once sorted
                        const unsigned b4 = (bmask[j] >> (k0 + 4 * g)) & 0xfu;
                        if (b4) {
#pragma unroll
                            for (int h = 2 * g; h < 2 * g + 2; h++) {
                                const int k = k0 + 2 * h;
                                if ((bmask[j] >> k) & 1u) {
                                    if (cur >= 0 && rflush) racc[cur * rmul] += acc;
                                    __syncwarp(); // the next run may touch the same node from another lane
                                    acc = 0.0;
                                    cur = sKey[j * 32 + k];
                                }
                                acc = __fma_rn(w2[h].x, v2[h].x, acc);
                                if ((bmask[j] >> (k + 1)) & 1u) {
                                    if (cur >= 0 && rflush) racc[cur * rmul] += acc;
                                    __syncwarp();
                                    acc = 0.0;
                                    cur = sKey[j * 32 + k + 1];
                                }
                                acc = __fma_rn(w2[h].y, v2[h].y, acc);
                            }
                        } else {
#pragma unroll
                            for (int h = 2 * g; h < 2 * g + 2; h++) {
                                acc = __fma_rn(w2[h].x, v2[h].x, acc);
                                acc = __fma_rn(w2[h].y, v2[h].y, acc);
                            }
                        }
                    }
                }
            }
            if (cur >= 0 && rflush) racc[cur * rmul] += acc;
            __syncwarp();
        }
        // ---- add the warp tile to the global deposit and clear it; the mover sums N, Px, Py, Pz (KM:406-411) of the
        //      particles that went through the tile are the tile totals of Den, U, V, W (bilinear weights sum to 1) ----
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int k = lane; k < TD; k += 32) {
            const double v = tile[k];
            if (v != 0.0) {
                const int f = k / (NT * NT), r = k % (NT * NT);
                const int gi = ti0 + r / NT, gj = tj0 + r % NT;
                if (gi >= 0 && gj >= 0 && gi < m.ni && gj < m.nj) atomicAdd(a.dep + f * plane + (size_t)gi * m.nj + gj, v);
                tile[k] = 0.0;
                if (f == 0) s0 += v;
                else if (f == 1) s1 += v;
                else if (f == 2) s2 += v;
                else if (f == 3) s3 += v;
            }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) {
            double *ex = tile + TD;
            if (s0 != 0 || ex[0] != 0) {
                atomicAdd(&a.c->sums[0], s0); atomicAdd(&a.c->sums[1], s1); atomicAdd(&a.c->sums[2], s2);
                atomicAdd(&a.c->sums[3], s3); atomicAdd(&a.c->sums[4], ex[0]);
            }
            ex[0] = 0.0;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// tail kernel: fast-store particles appended since the last sort (injection) -- same arithmetic, global deposit
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fast_tail(const __grid_constant__ FastStepArgs a, unsigned long long first, unsigned long long n)
{
    const MeshDev &m = a.m;
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
// deferred kernel: the few particles k_fast_step did not finish -- everything that is not the common case goes through
// sf_move() here, so that the hot kernel holds no call and no rare-path registers
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fast_deferred(const __grid_constant__ FastStepArgs a)
{
    const MeshDev &m = a.m;
    double sN = 0, sPx = 0, sPy = 0, sPz = 0, sE = 0;
    unsigned long long nd = a.c->n_defer[a.mesh_id];
    if (nd > a.defer_cap) nd = a.defer_cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned nfall = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += stride) {
        const unsigned e = a.defer[k];
        const size_t q = e & ~SF_DEFER_DEPOSIT_ONLY;
        PState p;
        bool deposit = false;
        if (e & SF_DEFER_DEPOSIT_ONLY) {
            p.x = a.fs.x[q]; p.y = a.fs.y[q]; p.z = a.fs.z[q]; p.u = a.fs.u[q]; p.v = a.fs.v[q]; p.w = a.fs.w[q]; p.mpw = a.fs.mpw[q];
            p.li = (p.x - m.x0) / m.dhx;
            p.lj = (p.y - m.y0) / m.dhy;
            p.dt = 0;
            deposit = true;
            nfall++;
        } else {
            MoveAux aux;
            int st = SF_REMOVED;
            bool exact = true;
            double z0 = 0;
            long long w0bits = 0;
            if (!fast_load_move(a, m, q, p, aux, st, exact, z0, w0bits)) continue;
            fast_epilogue(a, m, q, st, exact, p, aux, z0, w0bits, make_int2(0, 0), deposit);
        }
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
    if (nfall) atomicAdd(&a.c->n_fallback, (unsigned long long)nfall);
}

// ---------------------------------------------------------------------------------------------------------
// cell sort + compaction (K3): counting sort by cell key, out of place
// ---------------------------------------------------------------------------------------------------------

// pass 1: key and rank of every particle; hist[key] = particles per cell.  Vacant slots get SF_KEY_NONE.
// Four particles per thread (blockDim apart, so every load stays coalesced): the twelve loads and then the four
// histogram atomics of a thread are in flight together instead of one dependent chain per particle.
#define SF_COUNT_ILP 4
__global__ void __launch_bounds__(256)
k_sort_count(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs fs, unsigned long long n, int ntj, unsigned *__restrict__ hist,
             unsigned *__restrict__ keys, unsigned *__restrict__ ranks, double dt_pred)
{
    // dt_pred != 0: the key is the cell of pos + vel*dt_pred, i.e. (up to the field kick) the cell the particle will be deposited into by the
    // step that follows this sort: that step then finds its batches grouped by NEW cell, and every later step sees an order one step fresher.
    // Order only: no result depends on it.
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * (blockDim.x * SF_COUNT_ILP) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double mpw[SF_COUNT_ILP], x[SF_COUNT_ILP], y[SF_COUNT_ILP];
#pragma unroll
    for (int k = 0; k < SF_COUNT_ILP; k++) {
        const unsigned long long q = q0 + (unsigned long long)k * blockDim.x;
        mpw[k] = sf_vacant();
        x[k] = y[k] = 0.0;
        if (q < n) {
            mpw[k] = fs.mpw[q]; x[k] = fs.x[q]; y[k] = fs.y[q];
            if (dt_pred != 0) { x[k] += fs.u[q] * dt_pred; y[k] += fs.v[q] * dt_pred; }
        }
    }
    unsigned key[SF_COUNT_ILP], grp[SF_COUNT_ILP], base[SF_COUNT_ILP];
#pragma unroll
    for (int k = 0; k < SF_COUNT_ILP; k++) {
        key[k] = SF_KEY_NONE;
        if (mpw[k] == mpw[k])
            key[k] = sf_cell_key(m, sf_div_exact(x[k] - m.x0, m.dhx, m.rdhx, m.fastdiv), sf_div_exact(y[k] - m.y0, m.dhy, m.rdhy, m.fastdiv), ntj);
        const unsigned act = __ballot_sync(0xffffffffu, key[k] != SF_KEY_NONE);
        grp[k] = 0;
        base[k] = 0;
        if (key[k] != SF_KEY_NONE) {
            grp[k] = __match_any_sync(act, key[k]);
            if (lane == __ffs(grp[k]) - 1) base[k] = atomicAdd(&hist[key[k]], (unsigned)__popc(grp[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < SF_COUNT_ILP; k++) {
        const unsigned long long q = q0 + (unsigned long long)k * blockDim.x;
        if (key[k] != SF_KEY_NONE) {
            const unsigned b = __shfl_sync(grp[k], base[k], __ffs(grp[k]) - 1);
            ranks[q] = b + __popc(grp[k] & ((1u << lane) - 1u));
        }
        if (q < n) keys[q] = key[k];
    }
}

// pass 1 for the particles k_fast_step<.., PREP> could not key itself: the ones it left to k_fast_deferred (keyed from the state that kernel stored, or
// vacant by now) and the unsorted tail behind the sorted prefix (injection, records that became normal particles again).  Same outputs as k_sort_count.
__global__ void __launch_bounds__(256)
k_sort_count_fix(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs fs, const unsigned *__restrict__ defer, unsigned long long n_defer,
                 unsigned long long tail_first, unsigned long long n, int ntj, unsigned *__restrict__ hist, unsigned *__restrict__ keys,
                 unsigned *__restrict__ ranks, double dt_pred)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long q;
    if (t < n_defer) {
        const unsigned e = defer[t];
        if (e & SF_DEFER_DEPOSIT_ONLY) return; // pushed, stored and keyed by the step kernel; only its deposit was deferred
        q = e;
    } else {
        q = tail_first + (t - n_defer);
        if (q >= n) return;
    }
    const double mpw = fs.mpw[q];
    unsigned key = SF_KEY_NONE;
    if (mpw == mpw) {
        double x = fs.x[q], y = fs.y[q];
        if (dt_pred != 0) { x += fs.u[q] * dt_pred; y += fs.v[q] * dt_pred; }
        key = sf_cell_key(m, sf_div_exact(x - m.x0, m.dhx, m.rdhx, m.fastdiv), sf_div_exact(y - m.y0, m.dhy, m.rdhy, m.fastdiv), ntj);
        ranks[q] = atomicAdd(&hist[key], 1u);
    }
    keys[q] = key;
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

// pass 3 in two halves (default): 3a writes only the inverse permutation (4 bytes per particle, scattered), 3b lets thread d of the
// OUTPUT gather its particle: the eight 8-byte streams are then read scattered (the input is still roughly in cell order, so the
// sectors are shared by neighbouring lanes through L1) and written fully coalesced, instead of written scattered
#define SF_SORT_ILP 4
__global__ void __launch_bounds__(256)
k_sort_invert(unsigned long long n, const unsigned *__restrict__ offs, const unsigned *__restrict__ keys, const unsigned *__restrict__ ranks,
              unsigned *__restrict__ inv)
{
    const unsigned long long q0 = (unsigned long long)blockIdx.x * (blockDim.x * SF_SORT_ILP) + threadIdx.x;
    unsigned key[SF_SORT_ILP], rank[SF_SORT_ILP], off[SF_SORT_ILP];
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++) {
        const unsigned long long q = q0 + (unsigned long long)k * blockDim.x;
        key[k] = q < n ? keys[q] : SF_KEY_NONE;
        rank[k] = (key[k] != SF_KEY_NONE) ? ranks[q] : 0u;
    }
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++) off[k] = (key[k] != SF_KEY_NONE) ? offs[key[k]] : 0u;
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++)
        if (key[k] != SF_KEY_NONE) inv[off[k] + rank[k]] = (unsigned)(q0 + (unsigned long long)k * blockDim.x);
}

__global__ void __launch_bounds__(256)
k_sort_gather(FastPtrs in, FastPtrs out, unsigned long long n_out, const unsigned *__restrict__ inv)
{
    const unsigned long long d0 = (unsigned long long)blockIdx.x * (blockDim.x * SF_SORT_ILP) + threadIdx.x;
    size_t q[SF_SORT_ILP];
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++) {
        const unsigned long long d = d0 + (unsigned long long)k * blockDim.x;
        q[k] = d < n_out ? inv[d] : 0;
    }
    double v[SF_SORT_ILP][7];
    int2 tag[SF_SORT_ILP];
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++) { // all gathers of the thread in flight together
        v[k][0] = in.x[q[k]]; v[k][1] = in.y[q[k]]; v[k][2] = in.z[q[k]]; v[k][3] = in.u[q[k]];
        v[k][4] = in.v[q[k]]; v[k][5] = in.w[q[k]]; v[k][6] = in.mpw[q[k]];
        tag[k] = in.tag[q[k]];
    }
#pragma unroll
    for (int k = 0; k < SF_SORT_ILP; k++) {
        const unsigned long long d = d0 + (unsigned long long)k * blockDim.x;
        if (d < n_out) {
            out.x[d] = v[k][0]; out.y[d] = v[k][1]; out.z[d] = v[k][2]; out.u[d] = v[k][3];
            out.v[d] = v[k][4]; out.w[d] = v[k][5]; out.mpw[d] = v[k][6];
            out.tag[d] = tag[k];
        }
    }
}

// work items: each tile's run [offs[tile*SF_TILE^2], offs[(tile+1)*SF_TILE^2]) cut into pieces of <= SF_ITEM_MAX particles
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
              unsigned long long n, RecPtrs rec, unsigned long long rec_first, unsigned long long rec_cap, StepCounters *__restrict__ c,
              unsigned *__restrict__ hist, int ntj)
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
        if (hist) atomicAdd(&hist[sf_cell_key(m, p.li, p.lj, ntj)], 1u); // streaming store: population per cell key
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
