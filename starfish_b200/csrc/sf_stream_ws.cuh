// sf_stream_ws.cuh -- warp-specialised form of the streaming step (sf_stream.cuh): same phases, same arithmetic, but the push
// and the deposit of consecutive chunks overlap inside one CTA.
//
//   P group (8 warps, 128 registers after setmaxnreg): phase 1 of chunk k+1 -- gather / kick / move / locate, output rank,
//            new-cell rank, pushed state parked in the chunk's stage slot.
//   D group (16 warps, 64 registers): phases 2-4 of chunk k -- scan, permutation, out-of-place store, run reduction, node
//            gather -> global deposit.
// One CTA per SM owns all 64 K registers; the two groups never share a CTA-wide barrier (named barriers per group), they hand
// chunks over through mbarriers: full[3] (TMA landed), pushed[2] (phase 1 done), freed[2] (stage / operand set consumed).
// Three stage buffers: chunk k being deposited, k+1 being pushed, k+2 in flight.
#pragma once
#include "sf_stream.cuh"

#define SFW_CHUNK 512
#define SFW_PT 256                       // push threads
#define SFW_DT 512                       // deposit threads
#define SFW_THREADS (SFW_PT + SFW_DT)
#define SFW_DWARPS (SFW_DT / 32)
#define SFW_ROW (SFW_CHUNK + 2)
#define SFW_STAGE_DOUBLES (8 * SFW_ROW)
#define SFW_NPIECE_MAX (SFS_NCELL + SFW_CHUNK / SFS_PIECE + 2)
#define SFW_P4_PARTS (SFW_DWARPS / 4)
#ifndef SFW_P1_UNROLL
#define SFW_P1_UNROLL 1
#endif
constexpr int sfw_p1_unroll = SFW_P1_UNROLL;
#ifndef SFW_PREGS
#define SFW_PREGS 128
#endif
#ifndef SFW_DREGS
#define SFW_DREGS 56
#endif
static_assert(SFW_PT * SFW_PREGS + SFW_DT * SFW_DREGS <= SFW_THREADS * 80, "setmaxnreg re-deals the registers the CTA was launched with (80 per thread)");

static_assert(SFW_DWARPS % 4 == 0 && SFS_SCAN_WARPS <= SFW_DWARPS, "deposit group layout");

#define SFW_OFF_AUX (3 * SFW_STAGE_DOUBLES)                    // [2 sets][3][SFW_ROW]
#define SFW_OFF_S (SFW_OFF_AUX + 2 * 3 * SFW_ROW)              // [SFS_MAXP][SFS_SROW]
#define SFW_OFF_END (SFW_OFF_S + SFS_MAXP * SFS_SROW)
#define SFW_U32_WORDS (2 * SFW_CHUNK + 2 * SFW_CHUNK + 4 * SFS_NCELL + 2 * SFS_NCELL) // pkN[2], pkO[2], cnt[2][2][NCELL], offN, baseO
#define SFW_U16_WORDS (2 * SFW_CHUNK + SFW_CHUNK + 3 * SFW_NPIECE_MAX + SFS_NCELL + 8) // flO[2], perm, pieces, ordN, round starts
#define SFW_SMEM_BYTES (SFW_OFF_END * 8 + SFW_U32_WORDS * 4 + SFW_U16_WORDS * 2 + 64)

__device__ __forceinline__ void sfw_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void sfw_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sfs_smem(bar)) : "memory");
}

__device__ __forceinline__ void sfw_issue(const FastPtrs &fs, const SDesc &d, double *stage, unsigned long long *bar)
{
    const unsigned long long b0 = d.begin & ~1ULL;
    const unsigned nal = (unsigned)((d.begin - b0) + d.count + 1) & ~1u;
    const unsigned bytes = nal * 8u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage was read / written through the generic proxy
    sfs_mbar_expect_tx(bar, 8u * bytes);
    const void *src[8] = {fs.x + b0, fs.y + b0, fs.z + b0, fs.u + b0, fs.v + b0, fs.w + b0, fs.mpw + b0, fs.tag + b0};
#pragma unroll
    for (int r = 0; r < 8; r++) sfs_tma_load(stage + r * SFW_ROW, src[r], bytes, bar);
}

__global__ void __launch_bounds__(SFW_THREADS, 1)
k_stream_ws(const __grid_constant__ StreamArgs a, const FastStepArgs *__restrict__ ga)
{
    extern __shared__ __align__(128) unsigned char sfw_raw[];
    double *stage = reinterpret_cast<double *>(sfw_raw);       // [3][8][SFW_ROW]
    double *aux2 = stage + SFW_OFF_AUX;                        // [2][3][SFW_ROW]
    double *S = stage + SFW_OFF_S;
    unsigned *pkN2 = reinterpret_cast<unsigned *>(stage + SFW_OFF_END); // [2][CHUNK]
    unsigned *pkO2 = pkN2 + 2 * SFW_CHUNK;                     // [2][CHUNK]
    unsigned *cnt2 = pkO2 + 2 * SFW_CHUNK;                     // [2 sets][cntN | cntO][NCELL]
    unsigned *offN = cnt2 + 4 * SFS_NCELL;
    unsigned *baseO = offN + SFS_NCELL;
    short *flO2 = reinterpret_cast<short *>(baseO + SFS_NCELL); // [2][CHUNK]
    unsigned short *perm = reinterpret_cast<unsigned short *>(flO2 + 2 * SFW_CHUNK);
    unsigned short *pcOrd = perm + SFW_CHUNK;
    unsigned short *pcStart = pcOrd + SFW_NPIECE_MAX;
    unsigned short *pcLen = pcStart + SFW_NPIECE_MAX;
    unsigned short *ordN = pcLen + SFW_NPIECE_MAX;
    unsigned short *rndStart = ordN + SFS_NCELL;
    __shared__ __align__(8) unsigned long long sFull[3], sPushed[2], sFreed[2];
    __shared__ __align__(16) SDesc sDesc[4];
    __shared__ double sSums[5];
    __shared__ int sNPieces, sNRows, sNFall, sBox[8];

    const int tid = threadIdx.x, lane = tid & 31;
    const MeshDev &m = a.b.m;
    const int ntj = a.b.ntj;

    if (tid == 0) {
        for (int k = 0; k < 3; k++) sfs_mbar_init(&sFull[k], 1);
        for (int k = 0; k < 2; k++) { sfs_mbar_init(&sPushed[k], SFW_PT / 32); sfs_mbar_init(&sFreed[k], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sNFall = 0;
        const SDesc d0 = sfs_get_desc(a, blockIdx.x), d1 = sfs_get_desc(a, blockIdx.x + gridDim.x);
        sDesc[0] = d0;
        sDesc[1] = d1;
        if (d0.count) sfw_issue(a.b.fs, d0, stage, &sFull[0]);
    }
    if (tid < 5) sSums[tid] = 0.0;
    if (tid < 8) sBox[tid] = (tid & 1) ? -1 : SFS_RC;
    for (int k = tid; k < 4 * SFS_NCELL; k += SFW_THREADS) cnt2[k] = 0;
    __syncthreads();

    if (tid < SFW_PT) {
        // =====================================================================================================
        // P group: phase 1
        // =====================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SFW_PREGS));
        const bool simple_ok = !m.has_b && !m.any_seg && a.b.dt > 0;
        for (unsigned j = 0;; j++) {
            const SDesc cur = sDesc[j & 3];
            if (cur.count == 0) break;
            const int set = j & 1;
            if (j >= 2) sfs_mbar_wait(&sFreed[set], ((j - 2) >> 1) & 1u); // chunk j-2 consumed: its operand set and stage are free
            if (tid == 0) {
                const SDesc nx = sDesc[(j + 1) & 3];
                if (nx.count) sfw_issue(a.b.fs, nx, stage + ((j + 1) % 3) * SFW_STAGE_DOUBLES, &sFull[(j + 1) % 3]);
                const unsigned long long idx = (unsigned long long)blockIdx.x + (unsigned long long)(j + 2) * gridDim.x;
                SDesc *dst = &sDesc[(j + 2) & 3];
                if (idx < a.max_items) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sfs_smem(dst)), "l"(a.b.items + idx) : "memory");
                } else {
                    dst->begin = 0; dst->count = 0; dst->tile = -1;
                }
            }
            double *st = stage + (j % 3) * SFW_STAGE_DOUBLES;
            double *aux = aux2 + set * 3 * SFW_ROW;
            unsigned *pkN = pkN2 + set * SFW_CHUNK, *pkO = pkO2 + set * SFW_CHUNK;
            short *flO = flO2 + set * SFW_CHUNK;
            unsigned *cntN = cnt2 + set * 2 * SFS_NCELL, *cntO = cntN + SFS_NCELL;
            const int lead = (int)(cur.begin & 1ULL);
            const bool tiled = cur.tile >= 0;
            const int ci0 = tiled ? (cur.tile / ntj) * SF_TILE - SF_HALO : 0;
            const int cj0 = tiled ? (cur.tile % ntj) * SF_TILE - SF_HALO : 0;
            sfs_mbar_wait(&sFull[j % 3], (j / 3) & 1u);
#pragma unroll sfw_p1_unroll
            for (int o = tid; o < SFW_CHUNK; o += SFW_PT) {
                if (o >= cur.count) { flO[o] = -2; pkN[o] = 0xffffffffu; continue; }
                const int s = lead + o;
                PState p;
                p.mpw = st[6 * SFW_ROW + s];
                if (p.mpw != p.mpw) { flO[o] = -2; pkN[o] = 0xffffffffu; continue; } // vacant slot
                p.x = st[0 * SFW_ROW + s]; p.y = st[1 * SFW_ROW + s]; p.z = st[2 * SFW_ROW + s];
                p.u = st[3 * SFW_ROW + s]; p.v = st[4 * SFW_ROW + s]; p.w = st[5 * SFW_ROW + s];
                p.li = sf_div_exact(p.x - m.x0, m.dhx, m.rdhx, m.fastdiv);
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
                if (simple_ok && sf_move_simple(m, a.b.qm, a.b.dt, p)) {
                    st[0 * SFW_ROW + s] = p.x; st[1 * SFW_ROW + s] = p.y; st[2 * SFW_ROW + s] = p.z;
                    st[3 * SFW_ROW + s] = p.u; st[4 * SFW_ROW + s] = p.v; st[5 * SFW_ROW + s] = p.w;
                } else {
                    fl = stream_general(ga, st, SFW_ROW, s, sSums);
                    if (fl == 3) {
                        p.x = st[0 * SFW_ROW + s]; p.y = st[1 * SFW_ROW + s];
                        p.u = st[3 * SFW_ROW + s]; p.v = st[4 * SFW_ROW + s]; p.w = st[5 * SFW_ROW + s];
                        p.li = sf_div_ieee(p.x - m.x0, m.dhx);
                        p.lj = sf_div_ieee(p.y - m.y0, m.dhy);
                    } else {
                        st[6 * SFW_ROW + s] = sf_vacant();
                    }
                }
                unsigned pn = 0xffffffffu;
                if (fl == 3) {
                    const int i = sf_j2i(p.li), jj = sf_j2i(p.lj);
                    const bool inside = i >= 0 && jj >= 0 && i < m.ni - 1 && jj < m.nj - 1; // F2D:253
                    const int ri = i - ci0, rj = jj - cj0;
                    if (tiled && inside && ri >= 0 && rj >= 0 && ri < SFS_RC && rj < SFS_RC) {
                        const int ln = ri * SFS_RC + rj;
                        pn = ((unsigned)ln << 16) | atomicAdd(&cntN[ln], 1u);
                        double di, dj;
                        sfs_offsets(m, p.li, p.lj, i, jj, di, dj);
                        aux[0 * SFW_ROW + s] = di;
                        aux[1 * SFW_ROW + s] = dj;
                        aux[2 * SFW_ROW + s] = p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w); // KM:412
                    } else {
                        stream_fallback(ga, st, SFW_ROW, s, sSums, a.hist_next);
                        atomicAdd(&sNFall, 1);
                    }
                }
                pkN[o] = pn;
            }
            __syncwarp();
            if (lane == 0) sfw_arrive(&sPushed[set]); // release: this warp's parked state, ranks and counters
            if (tid == 0) asm volatile("cp.async.wait_all;" ::: "memory"); // descriptor of chunk j+2
            sfw_bar(1, SFW_PT);
        }
    } else {
        // =====================================================================================================
        // D group: phases 2 - 4
        // =====================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SFW_DREGS));
        const int dtid = tid - SFW_PT, wid = dtid >> 5;
        const size_t plane = (size_t)m.ni * m.nj;
        double msum0 = 0, msum1 = 0; // mover sums, KM:406-413
        for (unsigned j = 0;; j++) {
            const SDesc cur = sDesc[j & 3];
            if (cur.count == 0) break;
            const int set = j & 1;
            double *st = stage + (j % 3) * SFW_STAGE_DOUBLES;
            const double *aux = aux2 + set * 3 * SFW_ROW;
            const unsigned *pkN = pkN2 + set * SFW_CHUNK, *pkO = pkO2 + set * SFW_CHUNK;
            const short *flO = flO2 + set * SFW_CHUNK;
            unsigned *cntN = cnt2 + set * 2 * SFS_NCELL, *cntO = cntN + SFS_NCELL;
            const int lead = (int)(cur.begin & 1ULL);
            const bool tiled = cur.tile >= 0;
            const int ci0 = tiled ? (cur.tile / ntj) * SF_TILE - SF_HALO : 0;
            const int cj0 = tiled ? (cur.tile % ntj) * SF_TILE - SF_HALO : 0;
            sfs_mbar_wait(&sPushed[set], (j >> 1) & 1u);

            // ---------------- phase 2a: offsets, pieces, global bookkeeping ----------------
            unsigned gbase = 0;
            if (wid < SFS_SCAN_WARPS) {
                unsigned mine = 0, base = 0;
#pragma unroll
                for (int g = 0; g < SFS_SCAN_WARPS; g++) {
                    const int c = g * 32 + lane;
                    const unsigned cn = c < SFS_NCELL ? cntN[c] : 0u;
                    const unsigned v = cn | (((cn + SFS_PIECE - 1) / SFS_PIECE) << 12) | ((cn ? 1u : 0u) << 22);
                    const unsigned t = __reduce_add_sync(0xffffffffu, v);
                    if (g < wid) base += t;
                    if (g == wid) mine = v;
                }
                unsigned ic = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, ic, d);
                    if (lane >= d) ic += y;
                }
                ic += base;
                const unsigned ex = ic - mine, cn = mine & 0xfffu, np = (mine >> 12) & 0x3ffu;
                const unsigned oc = ex & 0xfffu, op = (ex >> 12) & 0x3ffu, orow = ex >> 22;
                const int c = wid * 32 + lane;
                int bi0 = SFS_RC, bi1 = -1, bj0 = SFS_RC, bj1 = -1;
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
                    if (c == SFS_NCELL - 1) {
                        const unsigned rows = ic >> 22, pieces = (ic >> 12) & 0x3ffu;
                        sNPieces = (int)pieces;
                        sNRows = (int)rows;
                        rndStart[rows ? (rows + SFS_MAXP - 1) / SFS_MAXP : 1] = (unsigned short)pieces;
                        if (!rows) rndStart[0] = 0;
                    }
                }
                bi0 = __reduce_min_sync(0xffffffffu, bi0); bi1 = __reduce_max_sync(0xffffffffu, bi1);
                bj0 = __reduce_min_sync(0xffffffffu, bj0); bj1 = __reduce_max_sync(0xffffffffu, bj1);
                if (lane == 0 && bi1 >= 0) {
                    atomicMin(&sBox[set * 4 + 0], bi0); atomicMax(&sBox[set * 4 + 1], bi1);
                    atomicMin(&sBox[set * 4 + 2], bj0); atomicMax(&sBox[set * 4 + 3], bj1);
                }
            }
            if (tiled) {
                const int c = SFW_DT - 1 - dtid;
                if (c < SFS_NCELL) {
                    const unsigned no = cntO[c], nn = cntN[c];
                    if (no | nn) {
                        const int ci = ci0 + c / SFS_RC, cj = cj0 + c % SFS_RC;
                        const unsigned key = sfs_gkey(ci, cj, ntj);
                        if (no) gbase = atomicAdd(&a.cursor[key], no);
                        if (nn) {
                            atomicAdd(&a.hist_next[key], nn);
                            atomicAdd(a.b.dep + SFGPU_F_MPC * plane + (size_t)ci * m.nj + cj, (double)nn); // KM:1593
                        }
                    }
                }
            }
            sfw_bar(2, SFW_DT); // D1

            // ---------------- phase 2b: permutation into new-cell order; output bases ----------------
            for (int o = dtid; o < SFW_CHUNK; o += SFW_DT) {
                const unsigned pn = pkN[o];
                if (pn != 0xffffffffu) perm[offN[pn >> 16] + (pn & 0xffffu)] = (unsigned short)(lead + o);
            }
            for (int k = dtid; k < 2 * SFS_NCELL; k += SFW_DT) cntN[k] = 0; // both counters of this set: chunk j+2 starts from zero
            if (dtid < 4) sBox[(set ^ 1) * 4 + dtid] = (dtid & 1) ? -1 : SFS_RC;
            for (int pid = rndStart[0] + dtid; pid < rndStart[1]; pid += SFW_DT) {
                const unsigned pc = pcOrd[pid];
                if ((pc & 0x8000u) && (pid == rndStart[0] || pcOrd[pid - 1] != pc))
                    for (int k = 0; k < 32; k++) S[(pc & 0x7fffu) * SFS_SROW + k] = 0.0;
            }
            if (tiled) {
                const int c = SFW_DT - 1 - dtid;
                if (c < SFS_NCELL) baseO[c] = gbase;
            }
            sfw_bar(2, SFW_DT); // D2

            // ---------------- output: every live input particle takes one slot of its old cell's segment ----------------
            for (int o = dtid; o < SFW_CHUNK; o += SFW_DT) {
                const int lo = flO[o];
                if (lo == -2) continue;
                const int s = lead + o;
                const size_t slot = (lo >= 0) ? (size_t)baseO[lo] + pkO[o] : (size_t)pkO[o];
                a.out.x[slot] = st[0 * SFW_ROW + s]; a.out.y[slot] = st[1 * SFW_ROW + s]; a.out.z[slot] = st[2 * SFW_ROW + s];
                a.out.u[slot] = st[3 * SFW_ROW + s]; a.out.v[slot] = st[4 * SFW_ROW + s]; a.out.w[slot] = st[5 * SFW_ROW + s];
                a.out.mpw[slot] = st[6 * SFW_ROW + s];
                a.out.tag[slot] = reinterpret_cast<const int2 *>(st + 7 * SFW_ROW)[s];
            }

            const int nrows = sNRows;
            for (int r0 = 0, rnd = 0; r0 < nrows; r0 += SFS_MAXP, rnd++) {
                const int pbeg = rndStart[rnd], pend = rndStart[rnd + 1];
                if (rnd > 0) {
                    sfw_bar(2, SFW_DT);
                    for (int pid = pbeg + dtid; pid < pend; pid += SFW_DT) {
                        const unsigned pc = pcOrd[pid];
                        if ((pc & 0x8000u) && (pid == pbeg || pcOrd[pid - 1] != pc))
                            for (int k = 0; k < 32; k++) S[(pc & 0x7fffu) * SFS_SROW + k] = 0.0;
                    }
                    sfw_bar(2, SFW_DT);
                }
                // ---------------- phase 3: cell totals.  Half a warp per piece: lane (slot s of 4, group g of 4) ----------------
                {
                    const int half = lane >> 4, s4 = (lane >> 2) & 3, g = lane & 3;
                    const double *Dv = (g == 0) ? (aux + 2 * SFW_ROW) : (st + (2 + g) * SFW_ROW);
                    const bool h0 = (lane & 4) != 0, h1 = (lane & 8) != 0;
                    for (int pp = pbeg + 2 * wid; pp < pend; pp += 2 * SFW_DWARPS) {
                        const int pid = pp + half;
                        const bool have = pid < pend;
                        const int start = have ? pcStart[pid] : 0, end = have ? start + pcLen[pid] : 0;
                        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
                        for (int k = start + s4; k < end; k += 4) {
                            const int q = perm[k];
                            const double di = aux[q], dj = aux[SFW_ROW + q], mp = st[6 * SFW_ROW + q], vc = Dv[q];
                            const double t = mp * vc;
                            const double v1 = g == 0 ? mp : t, v2 = g == 0 ? vc : t * vc; // KM:1584-1590
                            const double ai = 1 - di, bj = 1 - dj;
                            const double b1 = bj * v1, d1 = dj * v1, b2 = bj * v2, d2 = dj * v2;
                            a0 = __fma_rn(ai, b1, a0); a1 = __fma_rn(ai, b2, a1);
                            a2 = __fma_rn(di, b1, a2); a3 = __fma_rn(di, b2, a3);
                            a4 = __fma_rn(di, d1, a4); a5 = __fma_rn(di, d2, a5);
                            a6 = __fma_rn(ai, d1, a6); a7 = __fma_rn(ai, d2, a7);
                        }
                        double k0 = h0 ? a4 : a0, k1 = h0 ? a5 : a1, k2 = h0 ? a6 : a2, k3 = h0 ? a7 : a3;
                        k0 += __shfl_xor_sync(0xffffffffu, h0 ? a0 : a4, 4);
                        k1 += __shfl_xor_sync(0xffffffffu, h0 ? a1 : a5, 4);
                        k2 += __shfl_xor_sync(0xffffffffu, h0 ? a2 : a6, 4);
                        k3 += __shfl_xor_sync(0xffffffffu, h0 ? a3 : a7, 4);
                        double m0 = h1 ? k2 : k0, m1 = h1 ? k3 : k1;
                        m0 += __shfl_xor_sync(0xffffffffu, h1 ? k0 : k2, 8);
                        m1 += __shfl_xor_sync(0xffffffffu, h1 ? k1 : k3, 8);
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
                            msum0 += m0;
                            msum1 += m1;
                        }
                    }
                }
                sfw_bar(2, SFW_DT); // D3: cell totals visible; after the last round the stage and the operand set are dead
                if (r0 + SFS_MAXP >= nrows && dtid == 0) sfw_arrive(&sFreed[set]);
                // ---------------- phase 4: node totals -> global deposit ----------------
                {
                    const int bi0 = sBox[set * 4 + 0], bi1 = sBox[set * 4 + 1], bj0 = sBox[set * 4 + 2], bj1 = sBox[set * 4 + 3];
                    const int unit = 2 * wid + (lane >> 4), hl = lane & 15;
                    const int f = unit & 7, part = unit >> 3;
                    const int nrow = bi1 - bi0 + 1;
                    const int ra = bi0 + (nrow * part) / SFW_P4_PARTS, rb = bi0 + (nrow * (part + 1)) / SFW_P4_PARTS;
                    const int nb = bj0 + hl;
                    const bool cell_ok = nb <= bj1, node_ok = nb <= bj1 + 1;
                    const int col = (f >= 4 ? 4 : 0) + (f == 0 ? 0 : (f <= 3 ? f : f - 3));
                    const bool fok = f < 7;
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
            if (nrows == 0) { // nothing went through shared memory: the stage still has to be released
                sfw_bar(2, SFW_DT);
                if (dtid == 0) sfw_arrive(&sFreed[set]);
            }
            sfw_bar(2, SFW_DT); // D4: tables of this chunk are dead
        }
        // ---- mover sums: lanes (half, node n, group g) -> sum over the four nodes and the two halves ----
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
            msum0 += __shfl_xor_sync(0xffffffffu, msum0, o);
            msum1 += __shfl_xor_sync(0xffffffffu, msum1, o);
        }
        if (lane < 4 && msum0 != 0) atomicAdd(&a.b.c->sums[lane], msum0); // N, Px, Py, Pz
        if (lane == 0 && msum1 != 0) atomicAdd(&a.b.c->sums[4], msum1);   // E
    }
    __syncthreads();
    if (tid < 5 && sSums[tid] != 0) atomicAdd(&a.b.c->sums[tid], sSums[tid]);
    if (tid == 0 && sNFall) atomicAdd(&a.b.c->n_fallback, (unsigned long long)sNFall);
}
