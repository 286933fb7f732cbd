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

#ifndef SF_FAST_PAIRS
#define SF_FAST_PAIRS 4 // (push warp, deposit warp) pairs per CTA of the tiled kernel: warps 0..P-1 push, warps P..2P-1 deposit
#endif
#define SF_FAST_THREADS (2 * SF_FAST_PAIRS * 32)
#ifndef SF_FAST_MIN_CTAS
#define SF_FAST_MIN_CTAS 2
#endif
#define SF_FHALO 1 // cells kept around the 8x8 tile: the 3x3-cell window around any cell of the tile stays inside
#define SF_FNT (SF_TILE + 2 * SF_FHALO + 1) // nodes per edge of the deposit warp's tile
#define SF_TILE_DOUBLES (SFGPU_NFIELDS * SF_FNT * SF_FNT)
#ifndef SF_INTERLEAVE
#define SF_INTERLEAVE 1 // the sort interleaves every work item eight ways (one run per deposit quad)
#endif
#define SF_EXTRA 6 // per-warp sums of particles that missed the tile: [1..5] = N, Px, Py, Pz, (E)
// deposit operands of one 32-particle batch, published by the push warp for the deposit quads (two buffers per pair):
//   per slot: weight rows [0,0,1-d,d,0,0] per axis (Ruyten-corrected, F2D:265-283), 7 values, packed cell; then a header
#define SF_BUF_DOUBLES (32 * (6 + 6 + 7) + 16 + 4)
#define SF_NOFFS (SF_TILE * SF_TILE + 4) // cell boundaries of the item's tile, relative to the item
#ifndef SF_EHALO
#define SF_EHALO 2 // cells around the tile whose E nodes are staged in shared memory (particles drift between sorts)
#endif
#define SF_ENT (SF_TILE + 2 * SF_EHALO + 1) // nodes per edge of the staged E tile
#define SF_ETILE_DOUBLES (2 * SF_ENT * SF_ENT + (2 * SF_ENT * SF_ENT) % 2)
// one pair: tile | extra (deposit) | extra (push) | 2 buffers | offs | E tile | 4 mbarriers | 2 fallback counters (+pad)
#define SF_PAIR_DOUBLES (SF_TILE_DOUBLES + 2 * SF_EXTRA + 2 * SF_BUF_DOUBLES + SF_NOFFS / 2 + SF_ETILE_DOUBLES + 4 + 2)
#define SF_PAIR_SMEM_BYTES (SF_PAIR_DOUBLES * 8)
#define SF_FAST_SMEM_BYTES (SF_FAST_PAIRS * SF_PAIR_SMEM_BYTES)
static_assert(SF_PAIR_SMEM_BYTES % 16 == 0 && SF_TILE_DOUBLES % 2 == 0 && SF_BUF_DOUBLES % 2 == 0, "128-bit shared-memory rows");

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
    const unsigned *offs;     // first sorted position of every cell key of the layout the work items describe
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

// everything that is not the common case (boundaries, B field, segments, removal), out of line so that its
// register needs do not inflate the hot loop.  Returns true when the particle still deposits in this step.
__device__ __noinline__ bool fast_general(const FastStepArgs *__restrict__ ga, unsigned long long q, PState *pp, double z0, long long w0bits)
{
    const FastStepArgs &a = *ga;
    MoveAux aux;
    bool exact = true, deposit = false;
    const GlobalFieldGather fg;
    PState p = *pp;
    const int st = sf_move(a.m, a.meshes, a.qm, a.charge, a.dt, false, p, aux, exact, fg);
    fast_epilogue(a, a.m, q, st, exact, p, aux, z0, w0bits, make_int2(0, 0), deposit);
    *pp = p;
    return deposit;
}

// a particle that deposits outside the warp tile (drifted since the last sort): global FP64 REDs for the fields,
// and its mover sums (KM:406-413) into the warp's shared-memory slots (CAS atomics: only lanes of this warp contend)
__device__ __noinline__ void fast_fallback(const MeshDev *mp, const PState *pp, double *dep, double *extra, int *nfall)
{
    const PState &p = *pp;
    deposit_global(*mp, p, dep);
    atomicAdd(extra + 1, p.mpw);
    atomicAdd(extra + 2, p.mpw * p.u);
    atomicAdd(extra + 3, p.mpw * p.v);
    atomicAdd(extra + 4, p.mpw * p.w);
    atomicAdd(extra + 5, p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w));
    atomicAdd(nfall, 1);
}

// ---------------------------------------------------------------------------------------------------------
// straight-line common case of sf_move(): one substep, no B field, no segments, the particle starts and ends
// strictly inside the mesh.  No branches, so the compiler interleaves the SF_PPT independent particles a lane
// carries (FP64 latency hiding).  Every expression is the one sf_move() evaluates, in the same order, so the two
// paths are bit-identical; `ok` false means "not the common case": the caller re-runs the particle through
// sf_move() from its original state.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sf_move_simple(const MeshDev &m, double qm, double dt, PState &p)
{
    const int ni = m.ni, nj = m.nj;
    const int i = sf_j2i(p.li), j = sf_j2i(p.lj);
    bool ok = (p.mpw > 0) && i >= 0 && j >= 0 && i < ni - 1 && j < nj - 1;
    const int ic = min(max(i, 0), ni - 2), jc = min(max(j, 0), nj - 2); // safe addresses when !ok
    const double di = p.li - i, dj = p.lj - j;
    const size_t n00 = (size_t)ic * nj + jc;
    const double *fi = m.efi + n00, *fj = m.efj + n00;
    const double w00 = (1 - di) * (1 - dj), w10 = di * (1 - dj), w11 = di * dj, w01 = (1 - di) * dj; // F2D:345-348
    double ex = w00 * __ldg(fi);
    ex += w10 * __ldg(fi + nj);
    ex += w11 * __ldg(fi + nj + 1);
    ex += w01 * __ldg(fi + 1);
    double ey = w00 * __ldg(fj);
    ey += w10 * __ldg(fj + nj);
    ey += w11 * __ldg(fj + nj + 1);
    ey += w01 * __ldg(fj + 1);
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
    if (ok) {
        p.x = x; p.y = y; p.z = z; p.u = un; p.v = vn; p.w = wn; p.li = li; p.lj = lj;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------------------
// deposit of the tiled kernel.  FP64 shared-memory atomics are CAS loops on sm_100a and a transposition of the batch
// through shared memory ((node, field) lanes walking the particles) is bound by its operand loads, so the sums are
// kept in REGISTERS: four lanes (a "quad") own a run of consecutive particles of the cell-sorted work item (the sort
// interleaves the item eight ways, so lane l of every coalesced 32-particle batch belongs to run l & 7); lane a of
// the quad accumulates node row a of a 4x4-node window (3x3 cells around the run's current cell) for the 7 bilinear
// fields: 28 independent FMA chains, no atomics, no shuffles, no divergence.  A particle outside the window (the run
// reached the next cell) flushes the quad's registers into the warp-private tile with plain read-modify-writes and
// re-centres the window; quads flush one at a time.  Zero weights place a particle's 2x2 footprint inside the window.
// ---------------------------------------------------------------------------------------------------------
struct QuadAcc {
    double s[4][7];  // [node column b of the window][field]
    unsigned cnt;    // particles per cell (KM:1593) of window row a: three 10-bit counters (<= 256 particles per run)
    int bi, bj;      // window centre = the cell of the sorted layout the run is in, tile-array coordinates (cell = its lower-left node)
    int cur, t_next; // that cell (tile-local key) and the run-local index at which the next one begins
};

struct BatchScratch {
    double *wi, *wj;  // [32][6]: 0, 0, 1-d, d, 0, 0 -- a window lane reads its weight at a computed offset, no selects
    double2 *v01, *v23, *v45;
    double *v6;
    int *cell;
    const int *offs; // [SF_TILE*SF_TILE + 1] first sorted position of every cell of the tile, relative to the item
};

#define SF_PC_BIAS 64 // packed tile-array cell of a particle: ((ci + bias) << 8) | (cj + bias)

__device__ __forceinline__ void quad_flush(double *__restrict__ tile, int a, QuadAcc &A)
{
    double *row = tile + (A.bi - 1 + a) * SF_FNT + (A.bj - 1);
#pragma unroll
    for (int f = 0; f < 7; f++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            row[f * (SF_FNT * SF_FNT) + b] += A.s[b][f];
            A.s[b][f] = 0.0;
        }
    }
    if (a < 3 && A.cnt) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const unsigned n = (A.cnt >> (10 * c)) & 1023u;
            if (n) row[SFGPU_F_MPC * (SF_FNT * SF_FNT) + c] += (double)n;
        }
    }
    A.cnt = 0;
}

// the run of quad q reached run-local index t: move the window to the cell of the sorted layout that holds it
__device__ __forceinline__ void quad_advance(const BatchScratch &S, int q, int t, int L, QuadAcc &A)
{
    const int pos = (t < L) ? q * L + t : 8 * L + q; // the item's last count % 8 particles are not interleaved
    int cur = A.cur;
    while (cur < SF_TILE * SF_TILE - 1 && S.offs[cur + 1] <= pos) cur++;
    A.cur = cur;
    A.t_next = (t < L) ? min(S.offs[cur + 1] - q * L, L) : 0x7fffffff;
    if (cur == SF_TILE * SF_TILE - 1 && t < L) A.t_next = L;
    A.bi = cur / SF_TILE + SF_FHALO;
    A.bj = cur % SF_TILE + SF_FHALO;
}

// a particle outside its quad's window (it drifted more than a cell from where the last sort put it): global reductions
__device__ __noinline__ void quad_fallback(const FastStepArgs *__restrict__ ga, int a, int gi, int gj, double wx, double wj0, double wj1,
                                           const BatchScratch *Sp, int slot, double *extra, int *nfall)
{
    const double2 v01 = Sp->v01[slot], v23 = Sp->v23[slot], v45 = Sp->v45[slot];
    const double v[7] = {v01.x, v01.y, v23.x, v23.y, v45.x, v45.y, Sp->v6[slot]};
    const FastStepArgs &g = *ga;
    const size_t plane = (size_t)g.m.ni * g.m.nj;
    if (a < 2) { // node rows gi, gi + 1
        double *b = g.dep + (size_t)(gi + a) * g.m.nj + gj;
#pragma unroll
        for (int f = 0; f < 7; f++) {
            const double t = wx * v[f];
            atomicAdd(b + f * plane, wj0 * t);
            atomicAdd(b + f * plane + 1, wj1 * t);
        }
    } else if (a == 2) {
        atomicAdd(g.dep + SFGPU_F_MPC * plane + (size_t)gi * g.m.nj + gj, 1.0);
        atomicAdd(nfall, 1);
    } else {
        atomicAdd(extra + 1, v[0]);
        atomicAdd(extra + 2, v[1]);
        atomicAdd(extra + 3, v[2]);
        atomicAdd(extra + 4, v[3]);
    }
}

// one particle per quad: slot = the push lane that published it, t = its run-local index
__device__ __forceinline__ void quad_substep(const FastStepArgs *__restrict__ ga, double *__restrict__ tile, double *extra, int *nfall,
                                             const BatchScratch &S, int slot, int a, int q, int t, int L, int ti0, int tj0, QuadAcc &A)
{
    unsigned need = __ballot_sync(0xffffffffu, t >= A.t_next);
    while (need) { // quads flush one at a time: their windows may overlap in the warp tile
        const int ql = (__ffs(need) - 1) >> 2;
        if (q == ql) {
            quad_flush(tile, a, A);
            quad_advance(S, q, t, L, A);
        }
        __syncwarp();
        need &= ~(0xfu << (4 * ql));
    }
    const int pc = S.cell[slot];
    const int r = (pc >> 8) - (A.bi + SF_PC_BIAS - 1), c = (pc & 255) - (A.bj + SF_PC_BIAS - 1);
    const bool outside = pc >= 0 && ((unsigned)r > 2u || (unsigned)c > 2u);
    const double2 v01 = S.v01[slot], v23 = S.v23[slot], v45 = S.v45[slot];
    const double v[7] = {v01.x, v01.y, v23.x, v23.y, v45.x, v45.y, S.v6[slot]};
    if (__any_sync(0xffffffffu, outside)) {
        if (outside)
            quad_fallback(ga, a, ti0 + (pc >> 8) - SF_PC_BIAS, tj0 + (pc & 255) - SF_PC_BIAS, S.wi[6 * slot + 2 + (a & 1)], S.wj[6 * slot + 2],
                          S.wj[6 * slot + 3], &S, slot, extra, nfall);
        __syncwarp();
        if (outside && a == 0) S.wi[6 * slot + 2] = S.wi[6 * slot + 3] = 0.0; // nothing of it goes through the window
        __syncwarp();
    }
    // weights by offset into the padded rows: node row a of the window takes (1-di) when a == r, di when a == r+1, else 0
    const double wx = S.wi[6 * slot + min((unsigned)(2 - r + a), 5u)];
    const double *wyp = S.wj + 6 * slot + min((unsigned)(2 - c), 2u);
    const double wy0 = wyp[0], wy1 = wyp[1], wy2 = wyp[2], wy3 = wyp[3];
    if (a == r && pc >= 0 && !outside) A.cnt += 1u << (10 * c);
#pragma unroll
    for (int f = 0; f < 7; f++) {
        const double t_ = wx * v[f];
        A.s[0][f] = __fma_rn(wy0, t_, A.s[0][f]);
        A.s[1][f] = __fma_rn(wy1, t_, A.s[1][f]);
        A.s[2][f] = __fma_rn(wy2, t_, A.s[2][f]);
        A.s[3][f] = __fma_rn(wy3, t_, A.s[3][f]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// shared-memory barriers between the two warps of a pair (mbarrier, 32 arrivals per phase: every lane arrives after
// its own shared-memory accesses, every lane of the waiting warp acquires)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sf_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sf_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sf_saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sf_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sf_saddr(bar)) : "memory");
}
__device__ __forceinline__ void sf_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(sf_saddr(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

struct BatchHeader { // written by lane 0 of the push warp into every buffer it fills
    int kind;        // 0: a batch, 1: first batch of a work item, 2: no more work
    int count, tile; // of the work item
    int pad;
    unsigned long long begin;
};
static_assert(sizeof(BatchHeader) <= 4 * 8, "header slot");

struct PairSmem {
    double *tile, *extraD, *extraP, *buf[2], *etile;
    int *offs, *nfall; // nfall[0]: deposit warp, nfall[1]: push warp
    unsigned long long *full, *empty; // [2] each
};
__device__ __forceinline__ PairSmem sf_pair_smem(unsigned char *base)
{
    PairSmem P;
    P.tile = reinterpret_cast<double *>(base);
    P.extraD = P.tile + SF_TILE_DOUBLES;
    P.extraP = P.extraD + SF_EXTRA;
    P.buf[0] = P.extraP + SF_EXTRA;
    P.buf[1] = P.buf[0] + SF_BUF_DOUBLES;
    P.offs = reinterpret_cast<int *>(P.buf[1] + SF_BUF_DOUBLES);
    P.etile = reinterpret_cast<double *>(P.offs + SF_NOFFS);
    P.full = reinterpret_cast<unsigned long long *>(P.etile + SF_ETILE_DOUBLES);
    P.empty = P.full + 2;
    P.nfall = reinterpret_cast<int *>(P.empty + 2);
    return P;
}
__device__ __forceinline__ BatchScratch sf_buf_view(double *b, const int *offs)
{
    BatchScratch S;
    S.wi = b;
    S.wj = S.wi + 32 * 6;
    S.v01 = reinterpret_cast<double2 *>(S.wj + 32 * 6);
    S.v23 = S.v01 + 32;
    S.v45 = S.v23 + 32;
    S.v6 = reinterpret_cast<double *>(S.v45 + 32);
    S.cell = reinterpret_cast<int *>(S.v6 + 32);
    S.offs = offs;
    return S;
}
__device__ __forceinline__ BatchHeader *sf_buf_header(double *b) { return reinterpret_cast<BatchHeader *>(b + 32 * 19 + 16); }

// ---------------------------------------------------------------------------------------------------------
// push warp: gather (E tile in shared memory) -> kick -> move -> locate -> boundaries -> in-place store, one particle
// per lane; publishes every particle's deposit operands to the deposit warp of its pair
// ---------------------------------------------------------------------------------------------------------
template <int DOMAIN>
__device__ __forceinline__ void sf_push_role(const FastStepArgs &a, const FastStepArgs *__restrict__ ga, const PairSmem &P, int lane)
{
    const MeshDev &m = a.m;
    const bool simple_ok = !m.has_b && !m.any_seg && a.dt > 0;
    double *sE = P.etile;
    const unsigned n_items = *a.n_items;
    unsigned ph_empty[2] = {1u, 1u}; // a fresh barrier lets the first wait on "empty" pass
    int k = 0;
    double esum = 0.0; // sum of mpw*|vel| of this lane's particles (KM:412)
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(&a.c->queue[a.mesh_id], 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const WorkItem wi = a.items[it];
        const int ti0 = (wi.tile / a.ntj) * SF_TILE - SF_FHALO; // first node row / column of the deposit tile
        const int tj0 = (wi.tile % a.ntj) * SF_TILE - SF_FHALO;
        const int ei0 = ti0 + SF_FHALO - SF_EHALO, ej0 = tj0 + SF_FHALO - SF_EHALO; // first node of the staged E tile
        // E field of the tile's neighbourhood: the gathers of the common path read shared memory (F2D:300-350)
        __syncwarp();
        for (int e = lane; e < SF_ENT * SF_ENT; e += 32) {
            const int gi = ei0 + e / SF_ENT, gj = ej0 + e % SF_ENT;
            double fi = 0.0, fj = 0.0;
            if (gi >= 0 && gj >= 0 && gi < m.ni && gj < m.nj) {
                fi = __ldg(m.efi + (size_t)gi * m.nj + gj);
                fj = __ldg(m.efj + (size_t)gi * m.nj + gj);
            }
            sE[e] = fi;
            sE[SF_ENT * SF_ENT + e] = fj;
        }
        __syncwarp();
        // the state of the next batch is fetched into registers while the current one is pushed
        double nx = 0, ny = 0, nz = 0, nu = 0, nv = 0, nw = 0, nm = sf_vacant();
        if (lane < wi.count) {
            const size_t q0 = (size_t)wi.begin + lane;
            nx = a.fs.x[q0]; ny = a.fs.y[q0]; nz = a.fs.z[q0];
            nu = a.fs.u[q0]; nv = a.fs.v[q0]; nw = a.fs.w[q0];
            nm = a.fs.mpw[q0];
        }
        for (int b = 0; b < wi.count; b += 32) {
            PState p;
            const int o = b + lane;
            const bool present = o < wi.count;
            const size_t q = (size_t)wi.begin + o;
            p.x = nx; p.y = ny; p.z = nz; p.u = nu; p.v = nv; p.w = nw; p.mpw = nm;
            nm = sf_vacant();
            if (o + 32 < wi.count) {
                nx = a.fs.x[q + 32]; ny = a.fs.y[q + 32]; nz = a.fs.z[q + 32];
                nu = a.fs.u[q + 32]; nv = a.fs.v[q + 32]; nw = a.fs.w[q + 32];
                nm = a.fs.mpw[q + 32];
            }
            // ---- common case, straight line: one sub-step, no B field, no segments, the particle starts inside the staged E tile
            //      and ends strictly inside the mesh.  Every expression is the one sf_move() evaluates, in the same order (KM:336-381,
            //      F2D:345-348, UM:158-159), so both paths are bit-identical; anything else re-runs through sf_move() below ----
            p.li = sf_div_exact(p.x - m.x0, m.dhx, m.rdhx, m.fastdiv); // the stored lc of a normal particle is exactly XtoL(pos)
            p.lj = sf_div_exact(p.y - m.y0, m.dhy, m.rdhy, m.fastdiv);
            p.dt = 0;
            const int i0 = sf_j2i(p.li), j0 = sf_j2i(p.lj);
            const int ei = i0 - ei0, ej = j0 - ej0;
            bool done = present && simple_ok && p.mpw > 0 && (unsigned)ei < (unsigned)(SF_ENT - 1) && (unsigned)ej < (unsigned)(SF_ENT - 1) &&
                        (unsigned)i0 < (unsigned)(m.ni - 1) && (unsigned)j0 < (unsigned)(m.nj - 1);
            double xn, yn, zn, un, vn, wn, lin, ljn;
            {
                const double di = p.li - i0, dj = p.lj - j0;
                const double *e = sE + (done ? ei * SF_ENT + ej : 0);
                const double w00 = (1 - di) * (1 - dj), w10 = di * (1 - dj), w11 = di * dj, w01 = (1 - di) * dj; // F2D:345-348
                double ex = w00 * e[0];
                ex += w10 * e[SF_ENT];
                ex += w11 * e[SF_ENT + 1];
                ex += w01 * e[1];
                double ey = w00 * e[SF_ENT * SF_ENT];
                ey += w10 * e[SF_ENT * SF_ENT + SF_ENT];
                ey += w11 * e[SF_ENT * SF_ENT + SF_ENT + 1];
                ey += w01 * e[SF_ENT * SF_ENT + 1];
                un = p.u + a.qm * ex * a.dt; // KM:345-346 with part.dt = 0 + dt
                vn = p.v + a.qm * ey * a.dt;
                wn = p.w;
                xn = p.x + un * a.dt; // KM:369-370
                yn = p.y + vn * a.dt;
                if (DOMAIN == SFGPU_XY) {
                    zn = p.z + p.w * a.dt; // KM:380
                } else if (DOMAIN == SFGPU_RZ) { // KM:424-442
                    const double Aa = p.w * a.dt, Bb = xn, R = sqrt(Aa * Aa + Bb * Bb);
                    const double cs = Bb / R, sn = Aa / R;
                    zn = p.z - asin(sn);
                    xn = R;
                    const double u1 = un;
                    un = cs * u1 + sn * p.w;
                    wn = -sn * u1 + cs * p.w;
                } else { // KM:444-462
                    const double Aa = p.w * a.dt, Bb = yn, R = sqrt(Aa * Aa + Bb * Bb);
                    const double cs = Bb / R, sn = Aa / R;
                    zn = p.z + acos(cs);
                    yn = R;
                    const double v1 = vn;
                    vn = cs * v1 + sn * p.w;
                    wn = -sn * v1 + cs * p.w;
                }
                lin = sf_div_exact(xn - m.x0, m.dhx, m.rdhx, m.fastdiv); // UM:158-159
                ljn = sf_div_exact(yn - m.y0, m.dhy, m.rdhy, m.fastdiv);
                done = done && lin >= 0 && ljn >= 0 && lin < m.nim1 && ljn < m.njm1; // KM:606 (a NaN takes the general path)
            }
            bool deposit = false;
            if (done) { // alive, exact lc, dt == 0: in-place store
                a.fs.x[q] = xn;
                a.fs.y[q] = yn;
                if (zn != p.z) a.fs.z[q] = zn;
                a.fs.u[q] = un;
                a.fs.v[q] = vn;
                if (__double_as_longlong(wn) != __double_as_longlong(p.w)) a.fs.w[q] = wn;
                p.x = xn; p.y = yn; p.z = zn; p.u = un; p.v = vn; p.w = wn; p.li = lin; p.lj = ljn;
                deposit = true;
            } else if (present && p.mpw == p.mpw) { // general path (boundaries, B field, segments, removal, outside the E tile)
                PState t = p; // only the copy has its address taken: p stays in registers
                deposit = fast_general(ga, q, &t, p.z, __double_as_longlong(p.w));
                p = t;
            }
            // ---- deposit operands of this particle for its quad ----
            int pc = -1;
            double2 wi2 = make_double2(0.0, 0.0), wj2 = make_double2(0.0, 0.0);
            double val[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (deposit) {
                int ci, cj;
                double di, dj;
                const bool in = sf_deposit_axis_weights(m, p.li, p.lj, ci, cj, di, dj);
                const int li_ = ci - ti0, lj_ = cj - tj0;
                if (in && li_ >= -SF_PC_BIAS && lj_ >= -SF_PC_BIAS && li_ < 255 - SF_PC_BIAS && lj_ < 255 - SF_PC_BIAS) {
                    pc = ((li_ + SF_PC_BIAS) << 8) | (lj_ + SF_PC_BIAS);
                    wi2 = make_double2(1 - di, di);
                    wj2 = make_double2(1 - dj, dj);
                    sf_deposit_values(p, val);
                    esum += p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w); // KM:412
                } else {
                    const PState t = p;
                    fast_fallback(&ga->m, &t, a.dep, P.extraP, P.nfall + 1);
                }
            }
            // ---- hand the batch to the deposit warp ----
            sf_mbar_wait(P.empty + k, ph_empty[k]);
            ph_empty[k] ^= 1u;
            const BatchScratch S = sf_buf_view(P.buf[k], P.offs);
            S.cell[lane] = pc;
            *reinterpret_cast<double2 *>(S.wi + 6 * lane + 2) = wi2;
            *reinterpret_cast<double2 *>(S.wj + 6 * lane + 2) = wj2;
            S.v01[lane] = make_double2(val[0], val[1]);
            S.v23[lane] = make_double2(val[2], val[3]);
            S.v45[lane] = make_double2(val[4], val[5]);
            S.v6[lane] = val[6];
            if (lane == 0) {
                BatchHeader h;
                h.kind = b == 0 ? 1 : 0;
                h.count = wi.count;
                h.tile = wi.tile;
                h.pad = 0;
                h.begin = wi.begin;
                *sf_buf_header(P.buf[k]) = h;
            }
            sf_mbar_arrive(P.full + k);
            k ^= 1;
        }
    }
    // no more work: tell the deposit warp
    sf_mbar_wait(P.empty + k, ph_empty[k]);
    if (lane == 0) sf_buf_header(P.buf[k])->kind = 2;
    sf_mbar_arrive(P.full + k);
    // mover sums of this warp: energy of the particles it published, everything of those that missed the operand range
    esum = warp_sum(esum);
    __syncwarp();
    if (lane == 0) {
        const int nf = P.nfall[1];
        if (esum != 0 || nf != 0) {
            atomicAdd(&a.c->sums[0], P.extraP[1]); atomicAdd(&a.c->sums[1], P.extraP[2]); atomicAdd(&a.c->sums[2], P.extraP[3]);
            atomicAdd(&a.c->sums[3], P.extraP[4]); atomicAdd(&a.c->sums[4], esum + P.extraP[5]);
            if (nf != 0) atomicAdd(&a.c->n_fallback, (unsigned long long)nf);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// deposit warp: eight quads accumulate the published batches in registers, flush into the warp's tile, and add the
// tile to the global deposit at the end of every work item
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sf_deposit_item_end(const FastStepArgs &a, const PairSmem &P, int lane, int ti0, int tj0, QuadAcc &A)
{
    const MeshDev &m = a.m;
    const size_t plane = (size_t)m.ni * m.nj;
    const int qa = lane & 3, qq = lane >> 2;
    double *tile = P.tile;
#pragma unroll 1
    for (int ql = 0; ql < 8; ql++) { // every quad empties its registers into the warp tile, one quad at a time
        if (qq == ql) quad_flush(tile, qa, A);
        __syncwarp();
    }
    // the mover sums N, Px, Py, Pz (KM:406-411) of the particles that went through the tile are the tile totals of
    // Den, U, V, W (the weights of a particle sum to 1)
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int e = lane; e < SF_TILE_DOUBLES; e += 32) {
        const double v = tile[e];
        if (v != 0.0) {
            const int f = e / (SF_FNT * SF_FNT), r = e % (SF_FNT * SF_FNT);
            const int gi = ti0 + r / SF_FNT, gj = tj0 + r % SF_FNT;
            if (gi >= 0 && gj >= 0 && gi < m.ni && gj < m.nj) atomicAdd(a.dep + f * plane + (size_t)gi * m.nj + gj, v);
            tile[e] = 0.0;
            if (f == 0) s0 += v;
            else if (f == 1) s1 += v;
            else if (f == 2) s2 += v;
            else if (f == 3) s3 += v;
        }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == 0) {
        const int nf = P.nfall[0];
        if (s0 != 0 || nf != 0) {
            atomicAdd(&a.c->sums[0], s0 + P.extraD[1]); atomicAdd(&a.c->sums[1], s1 + P.extraD[2]); atomicAdd(&a.c->sums[2], s2 + P.extraD[3]);
            atomicAdd(&a.c->sums[3], s3 + P.extraD[4]);
            if (nf != 0) atomicAdd(&a.c->n_fallback, (unsigned long long)nf);
        }
#pragma unroll
        for (int e = 0; e < SF_EXTRA; e++) P.extraD[e] = 0.0;
        P.nfall[0] = 0;
    }
    __syncwarp();
}

__device__ __forceinline__ void sf_deposit_role(const FastStepArgs &a, const FastStepArgs *__restrict__ ga, const PairSmem &P, int lane)
{
    const int qa = lane & 3, qq = lane >> 2; // node row of the window, quad
    QuadAcc A;
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int f = 0; f < 7; f++) A.s[b][f] = 0.0;
    A.cnt = 0;
    A.bi = A.bj = 1;
    A.cur = 0;
    A.t_next = 0;
    unsigned ph_full[2] = {0u, 0u};
    int k = 0, b = 0, L = 0, ti0 = 0, tj0 = 0;
    bool have_item = false;
    for (;;) {
        sf_mbar_wait(P.full + k, ph_full[k]);
        ph_full[k] ^= 1u;
        const BatchHeader h = *sf_buf_header(P.buf[k]);
        if (h.kind != 0) {
            if (have_item) sf_deposit_item_end(a, P, lane, ti0, tj0, A);
            have_item = false;
            if (h.kind == 2) break;
            // a new work item: cell boundaries of the sorted layout inside it (positions relative to its first particle)
            ti0 = (h.tile / a.ntj) * SF_TILE - SF_FHALO;
            tj0 = (h.tile % a.ntj) * SF_TILE - SF_FHALO;
            L = h.count >> 3; // particles per run (sf_item_slot)
            for (int e = lane; e < SF_TILE * SF_TILE + 1; e += 32) {
                const long long d = (long long)a.offs[(size_t)h.tile * (SF_TILE * SF_TILE) + e] - (long long)h.begin;
                P.offs[e] = d < 0 ? 0 : (d > h.count ? h.count : (int)d);
            }
            __syncwarp();
            // first cell of every run: the number of boundaries at or before its first position
            {
                const int pos = qq * L;
                int n = 0;
                for (int e = qa; e < SF_TILE * SF_TILE; e += 4) n += (P.offs[e] <= pos) ? 1 : 0;
                n += __shfl_xor_sync(0xffffffffu, n, 1);
                n += __shfl_xor_sync(0xffffffffu, n, 2);
                A.cur = n > 0 ? n - 1 : 0;
            }
            A.t_next = 0; // the first sub-step places the window
            b = 0;
            have_item = true;
        }
        const BatchScratch S = sf_buf_view(P.buf[k], P.offs);
#pragma unroll 1
        for (int s = 0; s < 4; s++) quad_substep(ga, P.tile, P.extraD, P.nfall, S, 8 * s + qq, qa, qq, (b >> 3) + s, L, ti0, tj0, A);
        sf_mbar_arrive(P.empty + k);
        k ^= 1;
        b += 32;
    }
}

// ---------------------------------------------------------------------------------------------------------
// tiled kernel: persistent CTAs of SF_FAST_PAIRS (push warp, deposit warp) pairs; the push warp of a pair pulls work
// items from a queue and feeds its deposit warp batch by batch through two shared-memory buffers
// ---------------------------------------------------------------------------------------------------------
template <int DOMAIN> // SFGPU_XY / RZ / ZR: the rotation code of the axisymmetric movers stays out of the planar kernel
__global__ void __launch_bounds__(SF_FAST_THREADS, SF_FAST_MIN_CTAS)
k_fast_step(const __grid_constant__ FastStepArgs a, const FastStepArgs *__restrict__ ga)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pair = wid % SF_FAST_PAIRS;
    const bool pusher = wid < SF_FAST_PAIRS;
    const PairSmem P = sf_pair_smem(smem_raw + (size_t)pair * SF_PAIR_SMEM_BYTES);
    if (pusher) { // the pair's region starts zeroed (tile, sums, the padding of the weight rows); barriers: 32 arrivals
        for (int e = lane; e < SF_PAIR_DOUBLES; e += 32) P.tile[e] = 0.0;
        __syncwarp();
        if (lane < 2) {
            sf_mbar_init(P.full + lane, 32);
            sf_mbar_init(P.empty + lane, 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (pusher) sf_push_role<DOMAIN>(a, ga, P, lane);
    else sf_deposit_role(a, ga, P, lane);
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

// Work items: each tile's run [b, e) of the cell-sorted store is cut into `pieces` items of `per` particles (the last one
// shorter).  Inside an item of c particles the sort INTERLEAVES the cell order eight ways: with L = c / 8, sorted position
// p < 8L is stored at 8 (p % L) + p / L, so that lane l of every coalesced 32-particle batch of k_fast_step reads run l & 7
// of the item and each deposit quad sees a contiguous piece of the cell order; the c % 8 last particles keep their place.
struct ItemGeom {
    unsigned per, pieces;
};
__device__ __forceinline__ ItemGeom sf_item_geom(unsigned cnt)
{
    ItemGeom g;
    g.pieces = (cnt + SF_ITEM_MAX - 1) / SF_ITEM_MAX;
    g.per = ((cnt + g.pieces - 1) / g.pieces + 31u) & ~31u; // whole warps, balanced
    return g;
}
__device__ __forceinline__ unsigned sf_item_slot(unsigned b, unsigned e, unsigned d)
{
#if SF_INTERLEAVE
    const ItemGeom g = sf_item_geom(e - b);
    const unsigned k = (d - b) / g.per;
    const unsigned pb = b + k * g.per;
    const unsigned c = min(e, pb + g.per) - pb;
    const unsigned L = c >> 3, p = d - pb;
    if (p < 8u * L) {
        const unsigned r = p / L;
        return pb + 8u * (p - r * L) + r;
    }
#endif
    return d;
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
    const unsigned t0 = key & ~(unsigned)(SF_TILE * SF_TILE - 1);
    const size_t d = sf_item_slot(offs[t0], offs[t0 + SF_TILE * SF_TILE], offs[key] + ranks[q]);
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
    const ItemGeom g = sf_item_geom(e - b);
    const unsigned s = atomicAdd(n_items, g.pieces);
    for (unsigned k = 0; k < g.pieces && s + k < max_items; k++) {
        const unsigned pb = b + k * g.per, pe = min(e, pb + g.per);
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
