// sf_generic.cuh -- generic (untiled) kernels: one thread per full particle record, field gathers from
// global memory, deposit with FP64 global reductions (REDG.ADD.F64).  They are correct for ANY particle
// (stale lc, residual dt, mesh hand-off, slow path) and therefore serve three roles: the path for
// freshly injected / exceptional particles, the transfer sweeps of KM:131-142, and the cross-check of
// the tiled fast path (SFGPU_STEP_GENERIC).
#pragma once
#include "sf_device.cuh"
#include "sf_store.cuh"

__device__ __forceinline__ void rec_load(const RecPtrs &r, size_t q, PState &p, int2 &tag)
{
    p.x = r.x[q]; p.y = r.y[q]; p.z = r.z[q];
    p.u = r.u[q]; p.v = r.v[q]; p.w = r.w[q];
    p.mpw = r.mpw[q]; p.li = r.li[q]; p.lj = r.lj[q]; p.dt = r.dt[q];
    tag = r.tag[q];
}

__device__ __forceinline__ void rec_store(const RecPtrs &r, size_t q, const PState &p, int2 tag)
{
    r.x[q] = p.x; r.y[q] = p.y; r.z[q] = p.z;
    r.u[q] = p.u; r.v[q] = p.v; r.w[q] = p.w;
    r.mpw[q] = p.mpw; r.li[q] = p.li; r.lj[q] = p.lj; r.dt[q] = p.dt;
    r.tag[q] = tag;
}

// claim `want` consecutive slots for the lanes of this warp that have pred set; returns the lane's slot
__device__ __forceinline__ unsigned long long warp_claim(unsigned long long *cursor, bool pred)
{
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (mask) {
        const int leader = __ffs(mask) - 1;
        if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    return base + __popc(mask & ((1u << lane) - 1u));
}

// F2D:290-293 for the 7 bilinear fields + the cell count KM:1593, into the packed [8][ni][nj] buffer
__device__ __forceinline__ void deposit_global(const MeshDev &m, const PState &p, double *__restrict__ dep)
{
    const size_t plane = (size_t)m.ni * m.nj;
    DepW d;
    const bool in = sf_deposit_weights(m, p.li, p.lj, d);
    if (d.i >= 0 && d.j >= 0 && d.i < m.ni && d.j < m.nj)
        atomicAdd(dep + SFGPU_F_MPC * plane + (size_t)d.i * m.nj + d.j, 1.0);
    if (!in) return;
    double val[7];
    sf_deposit_values(p, val);
    const size_t n00 = (size_t)d.i * m.nj + d.j;
#pragma unroll
    for (int f = 0; f < 7; f++) {
        double *b = dep + f * plane + n00;
        atomicAdd(b, d.w00 * val[f]);
        atomicAdd(b + m.nj, d.w10 * val[f]);
        atomicAdd(b + m.nj + 1, d.w11 * val[f]);
        atomicAdd(b + 1, d.w01 * val[f]);
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ParticleMover.run over a list of full records (KM:298-422) fused with the deposit (KM:184-187, :1584-1593).
//  in[0..n_in)  -> survivors appended to out at *out_cursor, hand-offs to xfer[nb], slow path to slow.
__global__ void __launch_bounds__(256)
k_generic_step(const MeshDev *__restrict__ meshes, int mesh_id, double qm, double charge, double dt, int transfer,
               RecPtrs in, unsigned long long n_in, RecPtrs out, unsigned long long *__restrict__ out_cursor,
               unsigned long long out_cap, const XferDev *__restrict__ xfer, SlowPtrs slow, double *__restrict__ dep,
               StepCounters *__restrict__ c, FastPtrs fs, unsigned long long fast_cap, unsigned *__restrict__ hist, int ntj)
{
    const MeshDev m = meshes[mesh_id];
    const GlobalFieldGather fg;
    double sN = 0, sPx = 0, sPy = 0, sPz = 0, sE = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (n_in + stride - 1) / stride;
    for (unsigned long long r = 0; r < rounds; r++) {
        const unsigned long long q = r * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = q < n_in;
        PState p;
        int2 tag = make_int2(0, 0);
        MoveAux aux;
        int st = SF_REMOVED;
        bool exact = false;
        if (valid) {
            rec_load(in, q, p, tag);
            st = sf_move(m, meshes, qm, charge, dt, transfer != 0, p, aux, exact, fg);
        }
        const bool alive = valid && st == SF_ALIVE;
        if (alive) {
            deposit_global(m, p, dep);
            if (!transfer) { // KM:247: sums only on the non-transfer pass... (mover sums are discarded for transfers)
                sN += p.mpw;
                sPx += p.mpw * p.u;
                sPy += p.mpw * p.v;
                sPz += p.mpw * p.w;
                sE += p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w);
            }
        }
        // a survivor that is a normal particle again (lc == XtoL(pos), dt == 0) moves to the fast store's tail
        bool to_fast = alive && exact && p.dt == 0 && p.mpw == p.mpw && fast_cap > 0;
        const unsigned long long fslot = warp_claim(&c->fast_n[mesh_id], to_fast);
        if (to_fast && fslot >= fast_cap) to_fast = false;
        if (to_fast) {
            fs.x[fslot] = p.x; fs.y[fslot] = p.y; fs.z[fslot] = p.z;
            fs.u[fslot] = p.u; fs.v[fslot] = p.v; fs.w[fslot] = p.w;
            fs.mpw[fslot] = p.mpw;
            fs.tag[fslot] = tag;
            if (hist) { // streaming store: the per-cell population the next launch lays its output out by
                const int ci = min(max(sf_j2i(p.li), 0), m.ni - 2), cj = min(max(sf_j2i(p.lj), 0), m.nj - 2);
                atomicAdd(&hist[(unsigned)((ci / SF_TILE) * ntj + (cj / SF_TILE)) * (SF_TILE * SF_TILE) + (unsigned)((ci % SF_TILE) * SF_TILE + (cj % SF_TILE))], 1u);
            }
        }
        const unsigned nfast = __ballot_sync(0xffffffffu, to_fast);
        if ((threadIdx.x & 31) == 0 && nfast) atomicAdd((unsigned long long *)&c->fast_delta[mesh_id], (unsigned long long)__popc(nfast));
        const bool to_rec = alive && !to_fast;
        const unsigned long long slot = warp_claim(out_cursor, to_rec);
        if (to_rec) {
            if (slot < out_cap) rec_store(out, slot, p, tag);
            else atomicAdd(&c->overflow, 1ULL);
        }
        // mesh hand-off: copies into the neighbours' transfer lists (KM:715-721)
        const bool xf = valid && st == SF_TRANSFER;
        if (__any_sync(0xffffffffu, xf)) {
            if (xf) {
                for (int k = 0; k < 2; k++) {
                    if (!(aux.xfer_mask & (1 << k))) continue;
                    const int nb = aux.xfer_mesh[k];
                    const unsigned long long s = atomicAdd(&c->xfer_n[nb], 1ULL);
                    if (s < xfer[nb].cap) {
                        PState cp = p;
                        cp.li = aux.xfer_li[k];
                        cp.lj = aux.xfer_lj[k];
                        rec_store(xfer[nb].rec, s, cp, tag);
                    } else {
                        atomicAdd(&c->overflow, 1ULL);
                    }
                    atomicAdd(&c->n_xfer_copies, 1ULL);
                }
            }
        }
        const bool sl = valid && st == SF_SLOW;
        if (__any_sync(0xffffffffu, sl)) {
            if (sl) {
                const unsigned long long s = atomicAdd(&c->n_slow, 1ULL);
                if (s < slow.cap) {
                    rec_store(slow.rec, s, p, tag);
                    slow.old_x[s] = aux.xo; slow.old_y[s] = aux.yo;
                    slow.old_li[s] = aux.lio; slow.old_lj[s] = aux.ljo;
                    slow.bounces[s] = aux.bounces; slow.mesh[s] = mesh_id;
                } else {
                    atomicAdd(&c->overflow, 1ULL);
                }
            }
        }
        const unsigned absorbed = __ballot_sync(0xffffffffu, valid && st == SF_ABSORBED);
        if ((threadIdx.x & 31) == 0 && absorbed) atomicAdd(&c->n_absorbed, (unsigned long long)__popc(absorbed));
        const unsigned dead = __ballot_sync(0xffffffffu, valid && st == SF_DEAD);
        const unsigned rem = __ballot_sync(0xffffffffu, valid && st == SF_REMOVED);
        if ((threadIdx.x & 31) == 0) {
            if (dead) atomicAdd(&c->n_exited, (unsigned long long)__popc(dead));
            if (rem) atomicAdd(&c->n_removed, (unsigned long long)__popc(rem));
        }
    }
    if (!transfer) {
        sN = warp_sum(sN); sPx = warp_sum(sPx); sPy = warp_sum(sPy); sPz = warp_sum(sPz); sE = warp_sum(sE);
        if ((threadIdx.x & 31) == 0 && sN != 0) {
            atomicAdd(&c->sums[0], sN); atomicAdd(&c->sums[1], sPx); atomicAdd(&c->sums[2], sPy);
            atomicAdd(&c->sums[3], sPz); atomicAdd(&c->sums[4], sE);
        }
    }
}

// KineticMaterial.addParticle(MeshData, Particle), KM:759-802, in place over records [first, first+n)
__global__ void __launch_bounds__(256)
k_inject(const MeshDev *__restrict__ meshes, int mesh_id, double qm, double dt_step, int compute_lc, int rewind,
         RecPtrs r, unsigned long long first, unsigned long long n, StepCounters *__restrict__ c)
{
    const MeshDev m = meshes[mesh_id];
    const GlobalFieldGather fg;
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n) return;
    const size_t q = first + q0;
    PState p;
    int2 tag;
    rec_load(r, q, p, tag);
    if (compute_lc) { // KM:760-774
        p.li = (p.x - m.x0) / m.dhx;
        p.lj = (p.y - m.y0) / m.dhy;
        if (p.li >= m.ni) p.li = m.ni - 1;
        if (p.lj >= m.nj) p.lj = m.nj - 1;
    }
    if (rewind) { // KM:776-796
        sf_kick(m, qm, -0.5 * dt_step, p, fg);
        p.dt = 0;
    }
    if (!(isfinite(p.u) && isfinite(p.v) && isfinite(p.w))) atomicAdd(&c->n_bad, 1ULL); // KM:1357-1361
    r.u[q] = p.u; r.v[q] = p.v; r.w[q] = p.w;
    r.li[q] = p.li; r.lj[q] = p.lj; r.dt[q] = p.dt;
}

// deposit (+ mover sums) of records that were already moved this step on the host (slow-path survivors)
__global__ void __launch_bounds__(256)
k_deposit_records(const MeshDev *__restrict__ meshes, int mesh_id, RecPtrs r, unsigned long long first,
                  unsigned long long n, double *__restrict__ dep, StepCounters *__restrict__ c)
{
    const MeshDev m = meshes[mesh_id];
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    double sN = 0, sPx = 0, sPy = 0, sPz = 0, sE = 0;
    if (q0 < n) {
        PState p;
        int2 tag;
        rec_load(r, first + q0, p, tag);
        deposit_global(m, p, dep);
        sN = p.mpw; sPx = p.mpw * p.u; sPy = p.mpw * p.v; sPz = p.mpw * p.w;
        sE = p.mpw * sqrt(p.u * p.u + p.v * p.v + p.w * p.w);
    }
    sN = warp_sum(sN); sPx = warp_sum(sPx); sPy = warp_sum(sPy); sPz = warp_sum(sPz); sE = warp_sum(sE);
    if ((threadIdx.x & 31) == 0 && sN != 0) {
        atomicAdd(&c->sums[0], sN); atomicAdd(&c->sums[1], sPx); atomicAdd(&c->sums[2], sPy);
        atomicAdd(&c->sums[3], sPz); atomicAdd(&c->sums[4], sE);
    }
}

// KM:190-196: U,V,W /= Den (0 where Den==0, F2D:403-414); Den /= node_vol (F2D:418-431)
__global__ void k_moments(const double *__restrict__ dep, const double *__restrict__ node_vol, size_t plane,
                          double *__restrict__ out4)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= plane) return;
    const double den = dep[k];
    double u = dep[plane + k], v = dep[2 * plane + k], w = dep[3 * plane + k];
    if (den != 0) { u /= den; v /= den; w /= den; }
    else { u = 0; v = 0; w = 0; }
    out4[k] = den / node_vol[k];
    out4[plane + k] = u;
    out4[2 * plane + k] = v;
    out4[3 * plane + k] = w;
}

// restart.bin particle records (KineticMaterial.saveRestartData, KM:904-924): per particle, big-endian (DataOutputStream),
//   pos[0],vel[0],pos[1],vel[1],pos[2],vel[2], lc[0],lc[1], dt, mpw, mass (11 doubles), born_it, id (2 ints) = 96 bytes
#define SF_RESTART_RECORD_BYTES 96
__device__ __forceinline__ unsigned sf_bswap32(unsigned x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ unsigned long long sf_be64(double d)
{
    const unsigned long long v = (unsigned long long)__double_as_longlong(d);
    return ((unsigned long long)sf_bswap32((unsigned)v) << 32) | sf_bswap32((unsigned)(v >> 32));
}
__global__ void k_restart_pack(RecPtrs r, unsigned long long first, unsigned long long n, double mass, unsigned long long *__restrict__ out)
{
    const unsigned long long q0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n) return;
    const size_t q = first + q0;
    unsigned long long *o = out + q0 * (SF_RESTART_RECORD_BYTES / 8);
    o[0] = sf_be64(r.x[q]); o[1] = sf_be64(r.u[q]);
    o[2] = sf_be64(r.y[q]); o[3] = sf_be64(r.v[q]);
    o[4] = sf_be64(r.z[q]); o[5] = sf_be64(r.w[q]);
    o[6] = sf_be64(r.li[q]); o[7] = sf_be64(r.lj[q]);
    o[8] = sf_be64(r.dt[q]); o[9] = sf_be64(r.mpw[q]); o[10] = sf_be64(mass);
    const int2 tag = r.tag[q]; // {id, born_it}; the stream holds born_it first
    o[11] = ((unsigned long long)sf_bswap32((unsigned)tag.x) << 32) | sf_bswap32((unsigned)tag.y);
}

// running sums of updateSamples (KM:1584-1593): sums += this step's raw deposit
__global__ void k_accumulate(double *__restrict__ sums, const double *__restrict__ dep, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) sums[k] += dep[k];
}

// fails the context when the compiler contracted a*b+c (would break bit parity with Java)
__global__ void k_selftest_fmad(double a, double b, double c, double *out) { out[0] = a * b + c; }
