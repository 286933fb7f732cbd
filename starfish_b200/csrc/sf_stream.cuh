// sf_stream.cuh -- the streaming step: fused gather / push / locate / deposit that re-sorts the store as it writes.
//
// One launch does the move (KM:298-422), the deposit (KM:168-197, :1570-1595) and the cell sort + compaction (K3):
//   * the input store is cell-sorted as of the PREVIOUS step; a CTA takes a chunk of <= SFS_CHUNK particles of one
//     SF_TILE x SF_TILE-cell tile, fetched by TMA bulk copies (cp.async.bulk + mbarrier) one chunk ahead of the arithmetic;
//   * phase 1 (thread per particle): lc = XtoL(pos), E gather, kick, move, locate, boundaries -- the same expressions as
//     sf_move(), so results are bit-identical to the generic kernel;
//   * the particle is written OUT OF PLACE into the second slab at the segment of the cell it occupied BEFORE this push
//     (segment offsets = exclusive scan of the cell histogram the previous launch accumulated), so the output is
//     sorted by a cell that is one step old and holds no vacant slots: sort and compaction cost no extra pass;
//   * the deposit operands (corrected di, dj, mpw, u, v, w) are counting-sorted by NEW cell inside shared memory
//     (integer shared atomics give the rank), then warps walk cell runs: lane (slot s of 8, group g of 4) accumulates
//     the 4 node weights x 2 fields of every 8th particle in registers, a transposed shuffle reduction leaves one
//     (node, field) total per lane, and the per-run totals are stored -- no floating-point atomics in shared memory
//     (FP64 shared atomics are CAS loops on sm_100a);
//   * a last phase sums, for every node of the tile + halo, the totals of its four adjacent cells and issues one FP64
//     RED per touched (node, field) and chunk; the cell counts give mpc (KM:1593) and the next launch's histogram.
// Particles outside the tile + halo (fast movers, periodic wraps, freshly injected tail) use per-particle global
// atomics for all three purposes.  Exceptional particles leave for the record lists exactly as in sf_fast.cuh.
#pragma once
#include "sf_fast.cuh"

#ifndef SFS_CHUNK
#define SFS_CHUNK 1024  // particles per chunk (measured: 512 x 2 CTAs/SM is 6 % slower -- more runs and per-chunk overhead per particle)
#endif
#ifndef SFS_PPT
#define SFS_PPT 2       // particles per thread in phase 1
#endif
#define SFS_THREADS (SFS_CHUNK / SFS_PPT)
#define SFS_WARPS (SFS_THREADS / 32)
#ifndef SFS_MIN_CTAS
#define SFS_MIN_CTAS 1
#endif
#define SFS_RC (SF_TILE + 2 * SF_HALO)  // region cells per edge
#define SFS_NCELL (SFS_RC * SFS_RC)
#define SFS_NN (SFS_RC + 1)             // region nodes per edge
#ifndef SFS_PIECE
#define SFS_PIECE 64                    // longest run a warp reduces in one go
#endif
#define SFS_NPIECE_MAX (SFS_NCELL + SFS_CHUNK / SFS_PIECE + 2)
#define SFS_ROW (SFS_CHUNK + 2)         // stage row (doubles): an aligned superset of the chunk
#define SFS_NSTAGE 2
#ifndef SFS_P1_UNROLL
#define SFS_P1_UNROLL 1 // particles of a thread are pushed one after the other (interleaving them doubles the live registers)
#endif
constexpr int sfs_p1_unroll = SFS_P1_UNROLL;
#ifndef SFS_DYNAMIC
#define SFS_DYNAMIC 0   // 1: warps draw pieces from a shared counter instead of a fixed round robin (no gain measured)
#endif
#ifndef SFS_MAXP
#define SFS_MAXP 80                     // cell totals held per round (a chunk rarely fills more new cells)
#endif
#define SFS_SROW 33                     // run-total row (doubles), padded
#define SFS_STAGE_DOUBLES (8 * SFS_ROW)

static_assert(sizeof(WorkItem) == 16, "descriptors are fetched with one 16-byte cp.async");
static_assert(SFS_CHUNK < 0x1000 && SFS_NPIECE_MAX < 0x400 && SFS_NCELL < 0x400, "particles, pieces and cells are scanned as 12 + 10 + 10 bits");
static_assert(SFS_NCELL / SFS_MAXP + 2 <= 8, "round starts");
static_assert(SFS_NCELL <= SFS_THREADS - 32, "one thread per region cell besides warp 0");
static_assert(SFS_NN <= 16, "phase 4 gives half a warp to a node row");
#define SFS_SCAN_WARPS ((SFS_NCELL + 31) / 32)
static_assert(SFS_SCAN_WARPS <= SFS_WARPS, "one new cell per lane of the scanning warps");

struct StreamArgs {
    FastStepArgs b;        // b.fs = input store, b.items / b.n_items = chunks of the sorted prefix
    FastPtrs out;          // output store (second slab)
    unsigned *cursor;      // [nkeys] next free output slot of every cell key (starts at the segment offset)
    unsigned *hist_next;   // [nkeys] live particles per cell key after this step
    unsigned max_items;    // length of b.items (entries behind the last chunk have count 0)
};

struct SDesc {
    unsigned long long begin;
    int count;
    int tile;
};

__device__ __forceinline__ unsigned sfs_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sfs_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sfs_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sfs_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sfs_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sfs_tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sfs_smem(dst)), "l"(src),
                 "r"(bytes), "r"(sfs_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void sfs_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    if ((threadIdx.x & 31) == 0) // one lane polls; __syncwarp() orders the others behind the observed completion
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(sfs_smem(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
    __syncwarp();
}

// tile-major cell key of a (clamped) cell: the key of sf_cell_key()
__device__ __forceinline__ unsigned sfs_gkey(int ci, int cj, int ntj)
{
    return (unsigned)((ci / SF_TILE) * ntj + (cj / SF_TILE)) * (SF_TILE * SF_TILE) + (unsigned)((ci % SF_TILE) * SF_TILE + (cj % SF_TILE));
}

// what happens to a fast-store particle that does not stay a normal particle: record list, hand-off, slow path, death
__device__ __forceinline__ void fast_leave(const FastStepArgs &a, int st, const PState &p, const MoveAux &aux, int2 tag)
{
    atomicAdd((unsigned long long *)&a.c->fast_delta[a.mesh_id], ~0ULL); // -1
    if (st == SF_ALIVE) { // stale lc or residual dt: full record, handled by the generic kernel from now on
        const unsigned long long s = atomicAdd(&a.c->n_out[a.mesh_id], 1ULL);
        if (s < a.exc_cap) rec_store(a.exc, s, p, tag);
        else atomicAdd(&a.c->overflow, 1ULL);
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

// deposit of a particle the shared-memory path cannot take (outside the tile + halo, F2D:253 early return, exceptional):
// global FP64 REDs (F2D:290-293, KM:1593) and its mover sums (KM:406-413) into the CTA's shared slots (CAS atomics: rare)
__device__ __forceinline__ void sfs_deposit_global(const MeshDev &m, const PState &p, double *dep, double *sums)
{
    deposit_global(m, p, dep);
    atomicAdd(sums + 0, p.mpw);
    atomicAdd(sums + 1, p.mpw * p.u);
    atomicAdd(sums + 2, p.mpw * p.v);
    atomicAdd(sums + 3, p.mpw * p.w);
    atomicAdd(sums + 4, p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w));
}

// Everything that is not the common case, out of line.  The particle travels through its stage slot (shared memory), not
// through registers or a local struct: a call that keeps ten doubles alive costs the hot loop ~55 registers.
// Returns 3 when the particle stays a normal particle of the fast store (new state parked in the slot), else 0 (it left for a
// record list / died; if it still deposits in this step, that has been done here through the global path).
__device__ __noinline__ int stream_general(const FastStepArgs *__restrict__ ga, double *st, int row, int s, double *sums)
{
    const FastStepArgs &a = *ga;
    MoveAux aux;
    bool exact = true;
    const GlobalFieldGather fg;
    PState p;
    p.x = st[0 * row + s]; p.y = st[1 * row + s]; p.z = st[2 * row + s];
    p.u = st[3 * row + s]; p.v = st[4 * row + s]; p.w = st[5 * row + s];
    p.mpw = st[6 * row + s];
    p.li = (p.x - a.m.x0) / a.m.dhx; // the stored lc of a normal particle is exactly XtoL(pos)
    p.lj = (p.y - a.m.y0) / a.m.dhy;
    p.dt = 0;
    const int st_ = sf_move(a.m, a.meshes, a.qm, a.charge, a.dt, false, p, aux, exact, fg);
    st[0 * row + s] = p.x; st[1 * row + s] = p.y; st[2 * row + s] = p.z;
    st[3 * row + s] = p.u; st[4 * row + s] = p.v; st[5 * row + s] = p.w;
    if (st_ == SF_ALIVE && exact && p.dt == 0) return 3;
    fast_leave(a, st_, p, aux, reinterpret_cast<const int2 *>(st + 7 * row)[s]);
    if (st_ == SF_ALIVE) sfs_deposit_global(a.m, p, a.dep, sums); // stale lc / residual dt: it deposits where its lc says
    return 0;
}

// global-path deposit of a normal particle whose new cell the shared-memory path does not cover; counts it for the next launch
__device__ __noinline__ void stream_fallback(const FastStepArgs *__restrict__ ga, const double *st, int row, int s, double *sums, unsigned *hist_next)
{
    const MeshDev &m = ga->m;
    PState p;
    p.x = st[0 * row + s]; p.y = st[1 * row + s]; p.z = st[2 * row + s];
    p.u = st[3 * row + s]; p.v = st[4 * row + s]; p.w = st[5 * row + s];
    p.mpw = st[6 * row + s];
    p.li = (p.x - m.x0) / m.dhx;
    p.lj = (p.y - m.y0) / m.dhy;
    p.dt = 0;
    sfs_deposit_global(m, p, ga->dep, sums);
    atomicAdd(&hist_next[sf_cell_key(m, p.li, p.lj, ga->ntj)], 1u);
}

// the corrected in-cell offsets of F2D:262-288: the four weights are (1-di)(1-dj), di(1-dj), di*dj, (1-di)dj
__device__ __forceinline__ void sfs_offsets(const MeshDev &m, double fi, double fj, int i, int j, double &di, double &dj)
{
    di = fi - i;
    dj = fj - j;
    if (m.domain == SFGPU_RZ) {
        const double rp = m.x0 + (i + 1) * m.dhx, rm = m.x0 + i * m.dhx, r = m.x0 + fi * m.dhx;
        di = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    } else if (m.domain == SFGPU_ZR) {
        const double rp = m.y0 + (j + 1) * m.dhy, rm = m.y0 + j * m.dhy, r = m.y0 + fj * m.dhy;
        dj = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    }
}

// chunk descriptors come from the item list (tile chunks, then tail chunks; zero-filled behind the last one)
__device__ __forceinline__ SDesc sfs_get_desc(const StreamArgs &a, unsigned idx)
{
    SDesc d;
    d.begin = 0; d.count = 0; d.tile = -1;
    if (idx < a.max_items) {
        const WorkItem w = a.b.items[idx];
        d.begin = w.begin; d.count = w.count; d.tile = w.tile;
    }
    return d;
}

// 8 bulk copies (x,y,z,u,v,w,mpw,tag) of the 16-byte aligned superset of the chunk
__device__ __forceinline__ void sfs_issue(const FastPtrs &fs, const SDesc &d, double *stage, unsigned long long *bar)
{
    const unsigned long long b0 = d.begin & ~1ULL;
    const unsigned nal = (unsigned)((d.begin - b0) + d.count + 1) & ~1u;
    const unsigned bytes = nal * 8u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage was read / written through the generic proxy
    sfs_mbar_expect_tx(bar, 8u * bytes);
    const void *src[8] = {fs.x + b0, fs.y + b0, fs.z + b0, fs.u + b0, fs.v + b0, fs.w + b0, fs.mpw + b0, fs.tag + b0};
#pragma unroll
    for (int r = 0; r < 8; r++) sfs_tma_load(stage + r * SFS_ROW, src[r], bytes, bar);
}

// rows of S that several pieces add into (cells holding more than SFS_PIECE particles of the chunk) start from zero
__device__ __forceinline__ void sfs_zero_shared_rows(double *S, const unsigned short *pcOrd, int pbeg, int pend, int tid)
{
    for (int pid = pbeg + tid; pid < pend; pid += SFS_THREADS) {
        const unsigned pc = pcOrd[pid];
        if ((pc & 0x8000u) && (pid == pbeg || pcOrd[pid - 1] != pc))
            for (int k = 0; k < 32; k++) S[(pc & 0x7fffu) * SFS_SROW + k] = 0.0;
    }
}

#define SFS_P4_PARTS (SFS_WARPS / 4) // phase 4 deals 7 fields x SFS_P4_PARTS row ranges to the half warps
static_assert(SFS_WARPS % 4 == 0, "phase 4 deals 7 fields x row ranges to the half warps");
// shared memory of one CTA (dynamic), in doubles unless noted
#define SFS_OFF_AUX (SFS_NSTAGE * SFS_STAGE_DOUBLES)            // [3][SFS_ROW]: corrected di, dj, mpw*|vel| of the pushed particle
#define SFS_OFF_S (SFS_OFF_AUX + 3 * SFS_ROW)                   // [SFS_MAXP][SFS_SROW]: totals per non-empty new cell
#define SFS_OFF_END (SFS_OFF_S + SFS_MAXP * SFS_SROW)
#define SFS_U32_WORDS (2 * SFS_CHUNK + 4 * SFS_NCELL + 2 * SFS_NCELL) // pkN, pkO, cnt[2][2][NCELL], offN, baseO
#define SFS_U16_WORDS (2 * SFS_CHUNK + 3 * SFS_NPIECE_MAX + SFS_NCELL + 8) // flO, perm, pieces, ordN, round starts
#define SFS_SMEM_BYTES (SFS_OFF_END * 8 + SFS_U32_WORDS * 4 + SFS_U16_WORDS * 2 + 64)

__global__ void __launch_bounds__(SFS_THREADS, SFS_MIN_CTAS)
k_stream_step(const __grid_constant__ StreamArgs a, const FastStepArgs *__restrict__ ga)
{
    extern __shared__ __align__(128) unsigned char sfs_raw[];
    double *stage = reinterpret_cast<double *>(sfs_raw);       // [NSTAGE][8][SFS_ROW]: x,y,z,u,v,w,mpw,tag; phase 1 parks the pushed state in place
    double *aux = stage + SFS_OFF_AUX;
    double *S = stage + SFS_OFF_S;
    unsigned *pkN = reinterpret_cast<unsigned *>(stage + SFS_OFF_END); // [CHUNK] new region cell << 16 | rank in it, ~0: not in the shared-memory deposit
    unsigned *pkO = pkN + SFS_CHUNK;                           // [CHUNK] rank in the old region cell, or the absolute output slot
    unsigned *cnt = pkO + SFS_CHUNK;                           // [2 sets][cntN | cntO][NCELL], sets alternate between chunks
    unsigned *offN = cnt + 4 * SFS_NCELL;                      // [NCELL] first sorted position of a new cell
    unsigned *baseO = offN + SFS_NCELL;                        // [NCELL] first output slot of this chunk's share of an old cell's segment
    short *flO = reinterpret_cast<short *>(baseO + SFS_NCELL); // [CHUNK] old region cell, -1: pkO is the slot, -2: nothing to write
    unsigned short *perm = reinterpret_cast<unsigned short *>(flO + SFS_CHUNK); // [CHUNK] stage index of the k-th particle in new-cell order
    unsigned short *pcOrd = perm + SFS_CHUNK;                  // [NPIECE_MAX] pieces: row of S (bit 15: shared by several pieces), start, length
    unsigned short *pcStart = pcOrd + SFS_NPIECE_MAX;
    unsigned short *pcLen = pcStart + SFS_NPIECE_MAX;
    unsigned short *ordN = pcLen + SFS_NPIECE_MAX;             // [NCELL] 1 + ordinal of a non-empty new cell, 0: empty
    unsigned short *rndStart = ordN + SFS_NCELL;               // [<= 8] first piece of every round of SFS_MAXP cells
    __shared__ __align__(8) unsigned long long sBar[2];
    __shared__ __align__(16) SDesc sDesc[3];
    __shared__ double sSums[5];
    __shared__ int sNRows, sNFall, sBox[8]; // sBox: [2 sets][i min, i max, j min, j max]
#if SFS_DYNAMIC
    __shared__ int sNextPiece;
#endif

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const MeshDev &m = a.b.m;
    const size_t plane = (size_t)m.ni * m.nj;
    const bool simple_ok = !m.has_b && !m.any_seg && a.b.dt > 0;
    const int ntj = a.b.ntj;
    double msum0 = 0, msum1 = 0; // mover sums, KM:406-413: lane group g collects N, Px, Py, Pz (msum0) and E (msum1 of g = 0)

    // ---- chunks are dealt round robin; descriptor of chunk k+2 and data of chunk k+1 are in flight while k is processed ----
    if (tid == 0) {
        sfs_mbar_init(&sBar[0], 1);
        sfs_mbar_init(&sBar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sNFall = 0;
        const SDesc d0 = sfs_get_desc(a, blockIdx.x), d1 = sfs_get_desc(a, blockIdx.x + gridDim.x);
        sDesc[0] = d0;
        sDesc[1] = d1;
        if (d0.count) sfs_issue(a.b.fs, d0, stage, &sBar[0]);
    }
    if (tid < 5) sSums[tid] = 0.0;
    if (tid < 8) sBox[tid] = (tid & 1) ? -1 : SFS_RC;
    for (int k = tid; k < 4 * SFS_NCELL; k += SFS_THREADS) cnt[k] = 0;
    __syncthreads();

    for (unsigned it = 0;; it++) {
        const int sb = it & 1;
        const SDesc cur = sDesc[it % 3];
        if (cur.count == 0) break;
        if (tid == 0) {
            const SDesc nx = sDesc[(it + 1) % 3];
            if (nx.count) sfs_issue(a.b.fs, nx, stage + (sb ^ 1) * SFS_STAGE_DOUBLES, &sBar[sb ^ 1]);
            const unsigned long long idx = (unsigned long long)blockIdx.x + (unsigned long long)(it + 2) * gridDim.x;
            SDesc *dst = &sDesc[(it + 2) % 3];
            if (idx < a.max_items) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sfs_smem(dst)), "l"(a.b.items + idx) : "memory");
            } else {
                dst->begin = 0; dst->count = 0; dst->tile = -1;
            }
        }
        unsigned *cntN = cnt + sb * 2 * SFS_NCELL, *cntO = cntN + SFS_NCELL; // particles per new / old cell of this chunk
        double *st = stage + sb * SFS_STAGE_DOUBLES;
        const int lead = (int)(cur.begin & 1ULL);
        const bool tiled = cur.tile >= 0;
        const int ci0 = tiled ? (cur.tile / ntj) * SF_TILE - SF_HALO : 0; // first cell row / column of the region
        const int cj0 = tiled ? (cur.tile % ntj) * SF_TILE - SF_HALO : 0;
        sfs_mbar_wait(&sBar[sb], (it >> 1) & 1u);

        // ================= phase 1: push; the new state is parked in the particle's own stage slot =================
#pragma unroll sfs_p1_unroll
        for (int o = tid; o < SFS_CHUNK; o += SFS_THREADS) {
            if (o >= cur.count) { flO[o] = -2; pkN[o] = 0xffffffffu; continue; }
            const int s = lead + o;
            PState p;
            p.mpw = st[6 * SFS_ROW + s];
            if (p.mpw != p.mpw) { flO[o] = -2; pkN[o] = 0xffffffffu; continue; } // vacant slot (a particle that left during the previous step)
            p.x = st[0 * SFS_ROW + s]; p.y = st[1 * SFS_ROW + s]; p.z = st[2 * SFS_ROW + s];
            p.u = st[3 * SFS_ROW + s]; p.v = st[4 * SFS_ROW + s]; p.w = st[5 * SFS_ROW + s];
            p.li = sf_div_exact(p.x - m.x0, m.dhx, m.rdhx, m.fastdiv); // the stored lc of a normal particle is exactly XtoL(pos)
            p.lj = sf_div_exact(p.y - m.y0, m.dhy, m.rdhy, m.fastdiv);
            p.dt = 0;
            { // output segment: the cell the particle is in now
                const int ci = min(max(sf_j2i(p.li), 0), m.ni - 2), cj = min(max(sf_j2i(p.lj), 0), m.nj - 2);
                const int ri = ci - ci0, rj = cj - cj0;
                if (tiled && ri >= 0 && rj >= 0 && ri < SFS_RC && rj < SFS_RC) {
                    const int lo = ri * SFS_RC + rj;
                    flO[o] = (short)lo;
                    pkO[o] = atomicAdd(&cntO[lo], 1u);
                } else {
                    flO[o] = -1;
                    pkO[o] = atomicAdd(&a.cursor[sfs_gkey(ci, cj, ntj)], 1u);
                }
            }
            int fl = 3;
            if (simple_ok && sf_move_simple<false>(m, a.b.qm, a.b.dt, p, EGlobal())) {
                st[0 * SFS_ROW + s] = p.x; st[1 * SFS_ROW + s] = p.y; st[2 * SFS_ROW + s] = p.z;
                st[3 * SFS_ROW + s] = p.u; st[4 * SFS_ROW + s] = p.v; st[5 * SFS_ROW + s] = p.w;
            } else {
                fl = stream_general(ga, st, SFS_ROW, s, sSums);
                if (fl == 3) { // back to the common path: a normal particle again, lc = XtoL(pos)
                    p.x = st[0 * SFS_ROW + s]; p.y = st[1 * SFS_ROW + s];
                    p.u = st[3 * SFS_ROW + s]; p.v = st[4 * SFS_ROW + s]; p.w = st[5 * SFS_ROW + s];
                    p.li = sf_div_ieee(p.x - m.x0, m.dhx);
                    p.lj = sf_div_ieee(p.y - m.y0, m.dhy);
                } else {
                    st[6 * SFS_ROW + s] = sf_vacant(); // it left the fast store: the output slot becomes a vacant marker
                }
            }
            unsigned pn = 0xffffffffu;
            if (fl == 3) {
                const int i = sf_j2i(p.li), jj = sf_j2i(p.lj);
                const bool inside = i >= 0 && jj >= 0 && i < m.ni - 1 && jj < m.nj - 1; // F2D:253: scatter() returns early otherwise
                const int ri = i - ci0, rj = jj - cj0;
                if (tiled && inside && ri >= 0 && rj >= 0 && ri < SFS_RC && rj < SFS_RC) {
                    const int ln = ri * SFS_RC + rj;
                    pn = ((unsigned)ln << 16) | atomicAdd(&cntN[ln], 1u);
                    double di, dj;
                    sfs_offsets(m, p.li, p.lj, i, jj, di, dj);
                    aux[0 * SFS_ROW + s] = di;
                    aux[1 * SFS_ROW + s] = dj;
                    aux[2 * SFS_ROW + s] = p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w); // KM:412
                } else {
                    stream_fallback(ga, st, SFS_ROW, s, sSums, a.hist_next);
                    atomicAdd(&sNFall, 1);
                }
            }
            pkN[o] = pn;
        }
        __syncthreads(); // B1

        // ================= phase 2a: offsets, pieces, global bookkeeping =================
        unsigned gbase = 0;
        if (wid < SFS_SCAN_WARPS) { // one new cell per lane: particles, pieces and non-empty cells are scanned together (12 + 10 + 10 bits)
            unsigned mine = 0, base = 0;
#pragma unroll
            for (int g = 0; g < SFS_SCAN_WARPS; g++) { // every scanning warp sums all groups: no cross-warp hand-over, no extra barrier
                const int c = g * 32 + lane;
                const unsigned cn = c < SFS_NCELL ? cntN[c] : 0u;
                const unsigned v = cn | (((cn + SFS_PIECE - 1) / SFS_PIECE) << 12) | ((cn ? 1u : 0u) << 22);
                const unsigned t = __reduce_add_sync(0xffffffffu, v);
                if (g < wid) base += t;
                if (g == wid) mine = v;
            }
            unsigned ic = mine; // inclusive warp scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, ic, d);
                if (lane >= d) ic += y;
            }
            ic += base;
            const unsigned ex = ic - mine, cn = mine & 0xfffu, np = (mine >> 12) & 0x3ffu;
            const unsigned oc = ex & 0xfffu, op = (ex >> 12) & 0x3ffu, orow = ex >> 22;
            const int c = wid * 32 + lane;
            int bi0 = SFS_RC, bi1 = -1, bj0 = SFS_RC, bj1 = -1; // bounding box of the non-empty new cells
            if (c < SFS_NCELL) {
                offN[c] = oc;
                ordN[c] = (unsigned short)(cn ? orow + 1 : 0);
                if (cn) {
                    bi0 = bi1 = c / SFS_RC;
                    bj0 = bj1 = c % SFS_RC;
                    if (orow % SFS_MAXP == 0) rndStart[orow / SFS_MAXP] = (unsigned short)op;
                    const unsigned per = (cn + np - 1) / np;
                    for (unsigned q = 0; q < np; q++) {
                        const unsigned b = q * per, ln = min(per, cn - b);
                        pcOrd[op + q] = (unsigned short)((orow % SFS_MAXP) | (np > 1 ? 0x8000u : 0u));
                        pcStart[op + q] = (unsigned short)(oc + b);
                        pcLen[op + q] = (unsigned short)ln;
                    }
                }
                if (c == SFS_NCELL - 1) { // inclusive totals
                    const unsigned rows = ic >> 22, pieces = (ic >> 12) & 0x3ffu;
#if SFS_DYNAMIC
                    sNextPiece = 0;
#endif
                    sNRows = (int)rows;
                    rndStart[rows ? (rows + SFS_MAXP - 1) / SFS_MAXP : 1] = (unsigned short)pieces; // end of the last round
                    if (!rows) rndStart[0] = 0;
                }
            }
            bi0 = __reduce_min_sync(0xffffffffu, bi0); bi1 = __reduce_max_sync(0xffffffffu, bi1);
            bj0 = __reduce_min_sync(0xffffffffu, bj0); bj1 = __reduce_max_sync(0xffffffffu, bj1);
            if (lane == 0 && bi1 >= 0) {
                atomicMin(&sBox[sb * 4 + 0], bi0); atomicMax(&sBox[sb * 4 + 1], bi1);
                atomicMin(&sBox[sb * 4 + 2], bj0); atomicMax(&sBox[sb * 4 + 3], bj1);
            }
        }
        if (tiled) {
            const int c = SFS_THREADS - 1 - tid;
            if (c < SFS_NCELL) {
                const unsigned no = cntO[c], nn = cntN[c];
                if (no | nn) {
                    const int ci = ci0 + c / SFS_RC, cj = cj0 + c % SFS_RC;
                    const unsigned key = sfs_gkey(ci, cj, ntj);
                    if (no) gbase = atomicAdd(&a.cursor[key], no); // the result is needed for the output only
                    if (nn) {
                        atomicAdd(&a.hist_next[key], nn);
                        atomicAdd(a.b.dep + SFGPU_F_MPC * plane + (size_t)ci * m.nj + cj, (double)nn); // KM:1593
                    }
                }
            }
        }
        __syncthreads(); // B2

        // ================= phase 2b: permutation into new-cell order; output bases =================
#pragma unroll
        for (int j = 0; j < SFS_PPT; j++) {
            const int o = j * SFS_THREADS + tid;
            const unsigned pn = pkN[o];
            if (pn != 0xffffffffu) perm[offN[pn >> 16] + (pn & 0xffffu)] = (unsigned short)(lead + o);
        }
        if (tid < 4) sBox[(sb ^ 1) * 4 + tid] = (tid & 1) ? -1 : SFS_RC; // the next chunk's bounding box starts empty
        { // the other counter set is free (its chunk finished phase 4 before B1): clear it for the next chunk
            unsigned *nextc = cnt + ((it + 1) & 1) * 2 * SFS_NCELL;
            for (int k = tid; k < 2 * SFS_NCELL; k += SFS_THREADS) nextc[k] = 0;
        }
        sfs_zero_shared_rows(S, pcOrd, rndStart[0], rndStart[1], tid);
        if (wid != 0 && tiled) {
            const int c = SFS_THREADS - 1 - tid;
            if (c < SFS_NCELL) baseO[c] = gbase;
        }
        __syncthreads(); // B3

        // ================= output: every live input particle takes one slot of its old cell's segment =================
#pragma unroll
        for (int j = 0; j < SFS_PPT; j++) {
            const int o = j * SFS_THREADS + tid;
            const int lo = flO[o];
            if (lo == -2) continue;
            const int s = lead + o;
            const size_t slot = (lo >= 0) ? (size_t)baseO[lo] + pkO[o] : (size_t)pkO[o];
            a.out.x[slot] = st[0 * SFS_ROW + s]; a.out.y[slot] = st[1 * SFS_ROW + s]; a.out.z[slot] = st[2 * SFS_ROW + s];
            a.out.u[slot] = st[3 * SFS_ROW + s]; a.out.v[slot] = st[4 * SFS_ROW + s]; a.out.w[slot] = st[5 * SFS_ROW + s];
            a.out.mpw[slot] = st[6 * SFS_ROW + s];
            a.out.tag[slot] = reinterpret_cast<const int2 *>(st + 7 * SFS_ROW)[s];
        }

        const int nrows = sNRows;
        for (int r0 = 0, rnd = 0; r0 < nrows; r0 += SFS_MAXP, rnd++) {
            const int pbeg = rndStart[rnd], pend = rndStart[rnd + 1];
            if (rnd > 0) { // (round 0 was prepared before B3)
                __syncthreads(); // S is reused
#if SFS_DYNAMIC
                if (tid == 0) sNextPiece = 0;
#endif
                sfs_zero_shared_rows(S, pcOrd, pbeg, pend, tid);
                __syncthreads();
            }
            // ================= phase 3: cell totals.  Half a warp per piece: lane (slot s of 4, group g of 4) =================
            {
                const int half = lane >> 4, s4 = (lane >> 2) & 3, g = lane & 3;
                const double *Dv = (g == 0) ? (aux + 2 * SFS_ROW) : (st + (2 + g) * SFS_ROW); // g = 0: mpw*|vel|; g = 1..3: u, v, w
                const bool h0 = (lane & 4) != 0, h1 = (lane & 8) != 0;
#if SFS_DYNAMIC
                for (;;) { // warps draw pairs of pieces from a shared counter (runs differ in length)
                    int pp = 0;
                    if (lane == 0) pp = atomicAdd(&sNextPiece, 2);
                    pp = pbeg + __shfl_sync(0xffffffffu, pp, 0);
                    if (pp >= pend) break;
#else
                for (int pp = pbeg + 2 * wid; pp < pend; pp += 2 * SFS_WARPS) {
#endif
                    const int pid = pp + half;
                    const bool have = pid < pend;
                    const int start = have ? pcStart[pid] : 0, end = have ? start + pcLen[pid] : 0;
                    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0; // [node 00,10,11,01][field lo, hi]
                    for (int k = start + s4; k < end; k += 4) {
                        const int q = perm[k];
                        const double di = aux[q], dj = aux[SFS_ROW + q], mp = st[6 * SFS_ROW + q], vc = Dv[q];
                        const double t = mp * vc;
                        const double v1 = g == 0 ? mp : t, v2 = g == 0 ? vc : t * vc; // (Den | U,V,W) and (mpw*|vel| | UU,VV,WW), KM:1584-1590
                        const double ai = 1 - di, bj = 1 - dj;
                        const double b1 = bj * v1, d1 = dj * v1, b2 = bj * v2, d2 = dj * v2;
                        a0 = __fma_rn(ai, b1, a0); a1 = __fma_rn(ai, b2, a1); // (1-di)(1-dj)
                        a2 = __fma_rn(di, b1, a2); a3 = __fma_rn(di, b2, a3); // di(1-dj)
                        a4 = __fma_rn(di, d1, a4); a5 = __fma_rn(di, d2, a5); // di*dj
                        a6 = __fma_rn(ai, d1, a6); a7 = __fma_rn(ai, d2, a7); // (1-di)dj
                    }
                    // transposed reduction over the 4 slots: every stage halves the values a lane keeps
                    double k0 = h0 ? a4 : a0, k1 = h0 ? a5 : a1, k2 = h0 ? a6 : a2, k3 = h0 ? a7 : a3;
                    k0 += __shfl_xor_sync(0xffffffffu, h0 ? a0 : a4, 4);
                    k1 += __shfl_xor_sync(0xffffffffu, h0 ? a1 : a5, 4);
                    k2 += __shfl_xor_sync(0xffffffffu, h0 ? a2 : a6, 4);
                    k3 += __shfl_xor_sync(0xffffffffu, h0 ? a3 : a7, 4);
                    double m0 = h1 ? k2 : k0, m1 = h1 ? k3 : k1;
                    m0 += __shfl_xor_sync(0xffffffffu, h1 ? k0 : k2, 8);
                    m1 += __shfl_xor_sync(0xffffffffu, h1 ? k1 : k3, 8);
                    // this lane now holds node n = 2*h0 + h1 of group g: m0 = low field, m1 = high field
                    if (have) {
                        const unsigned pc = pcOrd[pid];
                        double *row = S + (pc & 0x7fffu) * SFS_SROW + ((h0 ? 4 : 0) + (h1 ? 2 : 0)) * 4 + g;
                        if (pc & 0x8000u) {
                            atomicAdd(row, m0);
                            atomicAdd(row + 4, m1);
                        } else {
                            row[0] = m0;
                            row[4] = m1;
                        }
                        msum0 += m0; // the node parts of a field add up to its particle total (the weights sum to 1)
                        msum1 += m1;
                    }
                }
            }
            __syncthreads(); // B4: cell totals visible
            // ================= phase 4: node totals -> global deposit =================
            // half a warp per field and half of the cell rows of the bounding box: a lane owns a node column, walks the cell rows
            // and hands the lower node row of each cell row to the next one in a register
            {
                const int bi0 = sBox[sb * 4 + 0], bi1 = sBox[sb * 4 + 1], bj0 = sBox[sb * 4 + 2], bj1 = sBox[sb * 4 + 3];
                const int unit = 2 * wid + (lane >> 4), hl = lane & 15;
                const int f = unit & 7, part = unit >> 3;
                const int nrow = bi1 - bi0 + 1;
                const int ra = bi0 + (nrow * part) / SFS_P4_PARTS, rb = bi0 + (nrow * (part + 1)) / SFS_P4_PARTS; // cell rows [ra, rb)
                const int nb = bj0 + hl; // node column; also the column of the cell whose node 00 / 10 it is
                const bool cell_ok = nb <= bj1, node_ok = nb <= bj1 + 1;
                const int col = (f >= 4 ? 4 : 0) + (f == 0 ? 0 : (f <= 3 ? f : f - 3));
                const bool fok = f < 7; // (the 16th unit idles; every lane still takes part in the shuffles)
                {
                    double carry = 0;
                    for (int ca = ra; ca < rb; ca++) {
                        double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
                        if (cell_ok && fok) {
                            const int ord = (int)ordN[ca * SFS_RC + nb] - 1;
                            if (ord >= r0 && ord < r0 + SFS_MAXP) {
                                const double *row = S + (ord - r0) * SFS_SROW + col;
                                t0 = row[0]; t1 = row[8]; t2 = row[16]; t3 = row[24];
                            }
                        }
                        const double u2 = __shfl_up_sync(0xffffffffu, t2, 1, 16), u3 = __shfl_up_sync(0xffffffffu, t3, 1, 16);
                        const double top = carry + (t0 + (hl ? u3 : 0.0));
                        carry = t1 + (hl ? u2 : 0.0);
                        if (fok && node_ok && top != 0.0) atomicAdd(a.b.dep + f * plane + (size_t)(ci0 + ca) * m.nj + (cj0 + nb), top);
                    }
                    if (fok && rb > ra && node_ok && carry != 0.0) atomicAdd(a.b.dep + f * plane + (size_t)(ci0 + rb) * m.nj + (cj0 + nb), carry);
                }
            }
        }
        if (nrows == 0) __syncthreads(); // (keeps the barrier count per chunk uniform: the stage must be dead before the next chunk's TMA)
        if (tid == 0) asm volatile("cp.async.wait_all;" ::: "memory"); // descriptor of chunk it+2 (published by the next barriers)
    }

    // ---- mover sums of this CTA: lanes (half, node n, group g) -> sum over the four nodes and the two halves ----
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
        msum0 += __shfl_xor_sync(0xffffffffu, msum0, o);
        msum1 += __shfl_xor_sync(0xffffffffu, msum1, o);
    }
    if (lane < 4 && msum0 != 0) atomicAdd(&a.b.c->sums[lane], msum0); // N, Px, Py, Pz
    if (lane == 0 && msum1 != 0) atomicAdd(&a.b.c->sums[4], msum1);   // E
    if (tid < 5 && sSums[tid] != 0) atomicAdd(&a.b.c->sums[tid], sSums[tid]);
    if (tid == 0 && sNFall) atomicAdd(&a.b.c->n_fallback, (unsigned long long)sNFall);
}


// chunks of the streaming kernel: each tile's run [offs[tile*SF_TILE^2], offs[(tile+1)*SF_TILE^2]) cut into <= SFS_CHUNK particles
__global__ void k_build_chunks(const unsigned *__restrict__ offs, int n_tiles, WorkItem *__restrict__ items, unsigned *__restrict__ n_items,
                               unsigned max_items, unsigned chunk)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const unsigned b = offs[(size_t)t * SF_TILE * SF_TILE], e = offs[(size_t)(t + 1) * SF_TILE * SF_TILE];
    if (e <= b) return;
    const unsigned cnt = e - b, pieces = (cnt + chunk - 1) / chunk;
    const unsigned s = atomicAdd(n_items, pieces);
    for (unsigned k = 0; k < pieces && s + k < max_items; k++) {
        // piece k = [b + k*cnt/pieces, b + (k+1)*cnt/pieces): never empty (pieces <= cnt), so a count of 0 can only be the
        // end-of-list marker the round-robin walk of k_stream_step stops at
        const unsigned pb = b + (unsigned)((unsigned long long)k * cnt / pieces), pe = b + (unsigned)((unsigned long long)(k + 1) * cnt / pieces);
        WorkItem w;
        w.begin = pb;
        w.count = (int)(pe - pb);
        w.tile = t;
        items[s + k] = w;
    }
}

// chunks of the unsorted tail (tile = -1: no region, everything through global atomics)
__global__ void k_build_tail(unsigned long long first, unsigned long long n, WorkItem *__restrict__ items, unsigned *__restrict__ n_items, unsigned max_items,
                             unsigned chunk)
{
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k * chunk >= n) return;
    const unsigned s = atomicAdd(n_items, 1u);
    if (s >= max_items) return;
    WorkItem w;
    w.begin = first + k * chunk;
    w.count = (int)((n - k * chunk) < chunk ? (n - k * chunk) : chunk);
    w.tile = -1;
    items[s] = w;
}

// ---------------------------------------------------------------------------------------------------------
// K3 as a streaming pass: re-sort of a store that is still ROUGHLY in cell order (it was sorted a few steps ago).  One chunk
// of one tile per CTA; a particle's destination = segment of its current cell (exclusive scan of the histogram) + this chunk's
// share of the segment (one global atomic per touched cell and chunk) + its rank among the chunk's particles of that cell
// (integer shared-memory atomic).  Reads are coalesced, writes leave in runs of same-cell particles: 128 B per particle at
// close to copy speed, where the generic counting sort scatters single 8-byte elements.
// ---------------------------------------------------------------------------------------------------------
#define SFR_THREADS 256
#define SFR_PPT 2
#define SFR_CHUNK (SFR_THREADS * SFR_PPT)
__global__ void __launch_bounds__(SFR_THREADS)
k_stream_sort(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs in, FastPtrs out, const WorkItem *__restrict__ items, unsigned max_items,
              unsigned *__restrict__ cursor, int ntj)
{
    __shared__ unsigned cntO[SFS_NCELL], baseO[SFS_NCELL];
    if (blockIdx.x >= max_items) return;
    const WorkItem it = items[blockIdx.x];
    if (it.count == 0) return;
    const MeshDev &m = meshes[mesh_id];
    const int tid = threadIdx.x;
    const bool tiled = it.tile >= 0;
    const int ci0 = tiled ? (it.tile / ntj) * SF_TILE - SF_HALO : 0, cj0 = tiled ? (it.tile % ntj) * SF_TILE - SF_HALO : 0;
    const double x0 = m.x0, y0 = m.y0, dhx = m.dhx, dhy = m.dhy;
    const int ni = m.ni, nj = m.nj;
    for (int k = tid; k < SFS_NCELL; k += SFR_THREADS) cntO[k] = 0;
    __syncthreads();
    double px[SFR_PPT], py[SFR_PPT], pz[SFR_PPT], pu[SFR_PPT], pv[SFR_PPT], pw[SFR_PPT], pm[SFR_PPT];
    int2 ptag[SFR_PPT];
    int lo[SFR_PPT];
    unsigned ro[SFR_PPT];
#pragma unroll
    for (int j = 0; j < SFR_PPT; j++) {
        const int o = j * SFR_THREADS + tid;
        lo[j] = -2;
        ro[j] = 0;
        if (o >= it.count) continue;
        const size_t q = (size_t)it.begin + o;
        pm[j] = in.mpw[q];
        if (pm[j] != pm[j]) continue; // vacant
        px[j] = in.x[q]; py[j] = in.y[q]; pz[j] = in.z[q]; pu[j] = in.u[q]; pv[j] = in.v[q]; pw[j] = in.w[q];
        ptag[j] = in.tag[q];
        const int ci = min(max(sf_j2i((px[j] - x0) / dhx), 0), ni - 2), cj = min(max(sf_j2i((py[j] - y0) / dhy), 0), nj - 2);
        const int ri = ci - ci0, rj = cj - cj0;
        if (tiled && ri >= 0 && rj >= 0 && ri < SFS_RC && rj < SFS_RC) lo[j] = ri * SFS_RC + rj;
        else {
            lo[j] = -1;
            ro[j] = atomicAdd(&cursor[sfs_gkey(ci, cj, ntj)], 1u);
        }
    }
#pragma unroll
    for (int j = 0; j < SFR_PPT; j++) { // ranks inside the region cells: one shared-memory atomic per distinct cell of the warp
        const unsigned act = __ballot_sync(0xffffffffu, lo[j] >= 0);
        if (lo[j] >= 0) {
            const unsigned grp = __match_any_sync(act, lo[j]);
            const int leader = __ffs(grp) - 1, lane = tid & 31;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&cntO[lo[j]], (unsigned)__popc(grp));
            ro[j] = __shfl_sync(grp, base, leader) + __popc(grp & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    for (int c = tid; c < SFS_NCELL; c += SFR_THREADS) {
        const unsigned no = cntO[c];
        if (no) baseO[c] = atomicAdd(&cursor[sfs_gkey(ci0 + c / SFS_RC, cj0 + c % SFS_RC, ntj)], no);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SFR_PPT; j++) {
        if (lo[j] == -2) continue;
        const size_t slot = lo[j] >= 0 ? (size_t)baseO[lo[j]] + ro[j] : (size_t)ro[j];
        out.x[slot] = px[j]; out.y[slot] = py[j]; out.z[slot] = pz[j];
        out.u[slot] = pu[j]; out.v[slot] = pv[j]; out.w[slot] = pw[j];
        out.mpw[slot] = pm[j];
        out.tag[slot] = ptag[j];
    }
}

// per-key histogram of the live particles of a store (re-establishes the streaming invariant after in-place edits)
__global__ void __launch_bounds__(256)
k_stream_hist(const MeshDev *__restrict__ meshes, int mesh_id, FastPtrs fs, unsigned long long first, unsigned long long n, int ntj,
              unsigned *__restrict__ hist)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned key = 0xffffffffu;
    if (q0 < n) {
        const size_t q = first + q0;
        const double mpw = fs.mpw[q];
        if (mpw == mpw) key = sf_cell_key(m, (fs.x[q] - m.x0) / m.dhx, (fs.y[q] - m.y0) / m.dhy, ntj);
    }
    // neighbours in a roughly sorted store share their cell: one atomic per distinct key of the warp
    const unsigned act = __ballot_sync(0xffffffffu, key != 0xffffffffu);
    if (key != 0xffffffffu) {
        const unsigned grp = __match_any_sync(act, key);
        if ((threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&hist[key], (unsigned)__popc(grp));
    }
}

// debug: every cell's cursor must have reached the start of the next segment
__global__ void k_stream_check(const unsigned *__restrict__ offs, const unsigned *__restrict__ cursor, unsigned nkeys, unsigned long long *bad)
{
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nkeys && cursor[k] != offs[k + 1]) atomicAdd(bad, 1ULL);
}
