// sf_device.cuh -- per-particle arithmetic of the Starfish kinetic hot path, sm_100a.
//
// Everything here must round exactly like the Java reference: the translation unit is compiled with
// -fmad=false (no mul+add contraction), FP64 division and sqrt are IEEE in CUDA, and (int)double is
// cvt.rzi.s32.f64 (toward zero, saturating, NaN -> 0) -- the same as Java's narrowing conversion.
// The only deliberate fused multiply-adds are the deposit accumulations, whose summation order is free
// anyway (BASELINE.json tolerance 1e-10).  sfgpu_create() runs a self test that fails if contraction
// was enabled by mistake.
//
// Java sources restated (src/starfish/core/): materials/KineticMaterial.java (KM),
// domain/Field2D.java (F2D), domain/UniformMesh.java (UM), domain/Mesh.java (MESH), common/Vec.java.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sfgpu.h"

#define SF_FLT_EPS 1e-7 // Constants.java:26
#define SF_MAX_BOUNCES 10 // KM:300

// per-particle outcome of a mover pass
enum : int {
    SF_ALIVE = 0,
    SF_REMOVED = 1,  // mpw <= 0, KM:322
    SF_DEAD = 2,     // OPEN / default face, CIRCUIT ion
    SF_SLOW = 3,     // host must run ProcessBoundary
    SF_TRANSFER = 4, // MESH hand-off, KM:708-722
    SF_ABSORBED = 5, // removed by the surface it hit, KM:586-603
};

// surface hits of a step (what the host needs for addSurfaceMomentum / addSurfaceMassDeposit / boundary_charge, KM:590-602)
struct HitList {
    int *seg, *mesh;
    double *t, *u, *v, *w, *mpw;
    signed char *alive;
    unsigned long long cap;
    unsigned long long *n; // cursor (device)
};

struct MeshDev {
    int ni, nj;
    int domain;  // SFGPU_XY / RZ / ZR
    int any_seg; // has_seg holds at least one 1
    int has_b;   // bfi/bfj present
    double x0, y0, dhx, dhy;
    double lenx, leny; // xd - x0 with xd = x0 + (n-1)*dh, UM:131-135, KM:700-706
    double nim1, njm1; // (double)(ni-1), (double)(nj-1): the KM:606 bounds without a conversion per particle
    double rdhx, rdhy; // RN(1/dh) for sf_div_exact
    int fastdiv;       // dh is a divisor sf_div_exact is proven for (host check in sfgpu_mesh_add)
    const int8_t *bc[4];
    const int *nbr[4];
    const uint8_t *has_seg;
    // node.segments restricted to DIRICHLET / SINK linear segments (sfgpu_mesh_set_segments): CSR over nodes i*nj+j.
    // Null: particles near a has_seg node are handed to the host (SF_SLOW) instead.
    const int *seg_offs, *seg_ids;
    const double4 *seg_xy; // x1, y1, x2, y2
    const int2 *seg_kind;  // {0: dies / 1: lives on unchanged, boundary is a SINK}
    HitList hits;
    int id;                // index of this mesh
    const double *efi, *efj, *bfi, *bfj;
    const double *node_vol;
};

struct PState {
    double x, y, z, u, v, w, mpw, li, lj, dt;
};

struct MoveAux {
    double xo, yo, lio, ljo; // pre-substep state (arguments of ProcessBoundary)
    int bounces;
    int xfer_mask;
    int xfer_mesh[2];
    double xfer_li[2], xfer_lj[2];
};

__device__ __forceinline__ int sf_j2i(double d) { return __double2int_rz(d); }

// XtoL's true division (UM:158-159) without the hardware division sequence.  With r = RN(1/b):
//   q0 = a*r is within 1.5 ulp of a/b; one residual correction q1 = q0 + (a - q0*b)*r is faithful; a second one
//   is the correctly rounded quotient (Markstein's theorem: r within 1/2 ulp of 1/b, q1 faithful, residual exact
//   through FMA).  The FMAs here are part of the division algorithm, not a contraction of reference arithmetic:
//   the result is bit-identical to IEEE a/b, which tools/div_check.c verifies on 1.2e9 dividends including
//   neighbours of rounding midpoints.  Outside the guarded range the IEEE division is used.
__device__ __noinline__ double sf_div_ieee(double a, double b) { return a / b; }
__device__ __forceinline__ double sf_div_exact(double a, double b, double r, bool fast)
{
    const double m = fabs(a);
    if (fast && m < 0x1p500 && m > 0x1p-500) {
        const double q0 = a * r;
        const double e0 = __fma_rn(-q0, b, a);
        const double q1 = __fma_rn(e0, r, q0);
        const double e1 = __fma_rn(-q1, b, a);
        return __fma_rn(e1, r, q1);
    }
    return sf_div_ieee(a, b);
}

// F2D:371-390
__device__ __forceinline__ double sf_gather_safe(const double *__restrict__ d, int ni, int nj, double fi, double fj)
{
    int i = sf_j2i(fi), j = sf_j2i(fj);
    double di = fi - i, dj = fj - j;
    if (i < 0) { i = 0; di = 0; }
    if (j < 0) { j = 0; dj = 0; }
    if (i >= ni - 1) { i = ni - 1; di = 0; }
    if (j >= nj - 1) { j = nj - 1; dj = 0; }
    const double *b = d + (size_t)i * nj + j;
    double v = (1 - di) * (1 - dj) * __ldg(b);
    if (di > 0) v += di * (1 - dj) * __ldg(b + nj);
    if (di > 0 && dj > 0) v += di * dj * __ldg(b + nj + 1);
    if (dj > 0) v += (1 - di) * dj * __ldg(b + 1);
    return v;
}

// F2D:300-350 (gather_safe when Java would have thrown IndexOutOfBounds)
__device__ __forceinline__ double sf_gather(const double *__restrict__ d, int ni, int nj, double fi, double fj)
{
    int i = sf_j2i(fi), j = sf_j2i(fj);
    if (i < 0 || j < 0 || i >= ni - 1 || j >= nj - 1) return sf_gather_safe(d, ni, nj, fi, fj);
    double di = fi - i, dj = fj - j;
    const double *b = d + (size_t)i * nj + j;
    double v = (1 - di) * (1 - dj) * __ldg(b);
    v += di * (1 - dj) * __ldg(b + nj);
    v += di * dj * __ldg(b + nj + 1);
    v += (1 - di) * dj * __ldg(b + 1);
    return v;
}

// gathers straight from global memory (L1/L2 cached); used by the generic kernels and as the
// out-of-tile fallback of the tiled kernel
struct GlobalFieldGather {
    __device__ __forceinline__ void operator()(const MeshDev &m, double li, double lj, double &ex, double &ey,
                                               double &bx, double &by) const
    {
        ex = sf_gather(m.efi, m.ni, m.nj, li, lj);
        ey = sf_gather(m.efj, m.ni, m.nj, li, lj);
        bx = 0;
        by = 0;
        if (m.has_b) {
            bx = sf_gather(m.bfi, m.ni, m.nj, li, lj);
            by = sf_gather(m.bfj, m.ni, m.nj, li, lj);
        }
    }
};

// KM:847-893 with E[2] = B[2] = 0 (ef/bf are fresh double[3], KM:303-304)
__device__ __forceinline__ void sf_boris(double qm, double dtp, double ex, double ey, double bx, double by, double &u,
                                         double &v, double &w)
{
    const double E[3] = {ex, ey, 0.0}, B[3] = {bx, by, 0.0};
    double vel[3] = {u, v, w};
    double t[3], s[3], vm[3], vp[3], vpl[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = qm * B[k] * 0.5 * dtp;
    double tm2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = 2 * t[k] / (1 + tm2);
#pragma unroll
    for (int k = 0; k < 3; k++) vm[k] = vel[k] + qm * E[k] * 0.5 * dtp;
    // Vec.CrossProduct3, Vec.java:318-325
    c[0] = vm[1] * t[2] - vm[2] * t[1];
    c[1] = -vm[0] * t[2] + vm[2] * t[0];
    c[2] = vm[0] * t[1] - vm[1] * t[0];
#pragma unroll
    for (int k = 0; k < 3; k++) vp[k] = vm[k] + c[k];
    c[0] = vp[1] * s[2] - vp[2] * s[1];
    c[1] = -vp[0] * s[2] + vp[2] * s[0];
    c[2] = vp[0] * s[1] - vp[1] * s[0];
#pragma unroll
    for (int k = 0; k < 3; k++) vpl[k] = vm[k] + c[k];
    u = vpl[0] + qm * E[0] * 0.5 * dtp;
    v = vpl[1] + qm * E[1] * 0.5 * dtp;
    w = vpl[2] + qm * E[2] * 0.5 * dtp;
}

// velocity update of KM:336-353 (also the rewind of KM:782-794 with dtp = -0.5*dt)
template <class FieldGather>
__device__ __forceinline__ void sf_kick(const MeshDev &m, double qm, double dtp, PState &p, const FieldGather &fg)
{
    double ex, ey, bx, by;
    fg(m, p.li, p.lj, ex, ey, bx, by);
    if (bx == 0 && by == 0) {
        p.u += qm * ex * dtp;
        p.v += qm * ey * dtp;
    } else {
        sf_boris(qm, dtp, ex, ey, bx, by, p.u, p.v, p.w);
    }
}

// Vec.mirror about a mesh-face normal, Vec.java:406-418 + UM:174-187
__device__ __forceinline__ void sf_mirror(PState &p, int face)
{
    double n0 = 0, n1 = 0, n2 = 0;
    if (face == SFGPU_FACE_LEFT) n0 = 1;
    else if (face == SFGPU_FACE_RIGHT) n0 = -1;
    else if (face == SFGPU_FACE_BOTTOM) n1 = 1;
    else n1 = -1;
    double tm = 0;
    tm += p.u * n0;
    tm += p.v * n1;
    tm += p.w * n2;
    double t0 = n0 * tm, t1 = n1 * tm, t2 = n2 * tm;
    double a0 = p.u - t0, a1 = p.v - t1, a2 = p.w - t2;
    t0 = t0 * -1;
    t1 = t1 * -1;
    t2 = t2 * -1;
    p.u = t0 + a0;
    p.v = t1 + a1;
    p.w = t2 + a2;
}

// Math.min / Math.max NaN semantics followed by (int), KM:484-492
__device__ __forceinline__ int sf_min_i(double a, double b) { return (a != a || b != b) ? 0 : sf_j2i(a < b ? a : b); }
__device__ __forceinline__ int sf_max_i(double a, double b) { return (a != a || b != b) ? 0 : sf_j2i(a > b ? a : b); }

// KM:482-518: is a DIRICHLET/SINK segment attached to a node of the substep's bounding box?
__device__ __forceinline__ bool sf_bbox_has_segments(const MeshDev &m, double li, double lj, double lio, double ljo)
{
    int i_min = sf_min_i(li, lio), i_max = sf_max_i(li, lio);
    int j_min = sf_min_i(lj, ljo), j_max = sf_max_i(lj, ljo);
    if (i_min < 0) i_min = 0;
    if (j_min < 0) j_min = 0;
    if (i_max >= m.ni) i_max = m.ni - 1;
    if (j_max >= m.nj) j_max = m.nj - 1;
    for (int i = i_min; i <= i_max; i++)
        for (int j = j_min; j <= j_max; j++)
            if (m.has_seg[(size_t)i * m.nj + j]) return true;
    return false;
}

// LinearSegment.intersect(p3, p4), LinearSegment.java:113-179: t0 along the segment, t1 along p3 -> p4; false if none
__device__ __forceinline__ bool sf_segment_intersect(const double4 sg, double x3, double y3, double x4, double y4, double &t0, double &t1)
{
    const double x1 = sg.x, y1 = sg.y, x2 = sg.z, y2 = sg.w;
    const double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (den == 0) return false;
    const double xp0 = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / den;
    const double xp1 = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / den;
    if (fabs(x2 - x1) > 1e-6) t0 = (xp0 - x1) / (x2 - x1);
    else t0 = (xp1 - y1) / (y2 - y1);
    if (t0 < -SF_FLT_EPS || t0 > (1 + SF_FLT_EPS)) return false;
    if (fabs(x4 - x3) > 1e-6) t1 = (xp0 - x3) / (x4 - x3);
    else t1 = (xp1 - y3) / (y4 - y3);
    if (t1 < -SF_FLT_EPS || t1 > (1 + SF_FLT_EPS)) return false;
    if (t0 < 0) t0 = 0;
    if (t1 < 0) t1 = 0;
    if (t0 > 1) t0 = 1;
    if (t1 > 1) t1 = 1;
    return true;
}

// segment part of ProcessBoundary, KM:482-603, for linear segments whose surface outcome is deterministic.
// 0: no hit; 1: hit, alive (p.x, p.y, p.li, p.lj, p.dt updated); 2: hit and removed.
__device__ __noinline__ int sf_process_segments(const MeshDev *mp, double dt0, double xo, double yo, double lio, double ljo, PState *pp)
{
    const MeshDev &m = *mp;
    PState &p = *pp;
    int i_min = sf_min_i(p.li, lio), i_max = sf_max_i(p.li, lio);
    int j_min = sf_min_i(p.lj, ljo), j_max = sf_max_i(p.lj, ljo);
    if (i_min < 0) i_min = 0;
    if (j_min < 0) j_min = 0;
    if (i_max >= m.ni) i_max = m.ni - 1;
    if (j_max >= m.nj) j_max = m.nj - 1;
    double tp_min = 2.0, tsurf_min = 0;
    int seg_min = -1;
    for (int i = i_min; i <= i_max; i++)
        for (int j = j_min; j <= j_max; j++) {
            const size_t node = (size_t)i * m.nj + j;
            for (int k = m.seg_offs[node]; k < m.seg_offs[node + 1]; k++) { // (a segment met twice gives the same t: no set needed)
                const int sid = m.seg_ids[k];
                const double4 sg = m.seg_xy[sid];
                double t0, t1;
                if (!sf_segment_intersect(sg, xo, yo, p.x, p.y, t0, t1)) continue;
                if (t1 > 0) { // KM:535
                    double dx = sg.z - sg.x, dy = sg.w - sg.y; // LinearSegment.java:26-44
                    const double len = sqrt(dx * dx + dy * dy);
                    dx /= len;
                    dy /= len;
                    const double acos_ = (-dy * p.u + dx * p.v) / sqrt(p.u * p.u + p.v * p.v);
                    if (t1 < SF_FLT_EPS && acos_ > 0) continue; // KM:541-544
                    if (t1 < tp_min) {
                        tp_min = t1;
                        tsurf_min = t0;
                        seg_min = sid;
                    }
                }
            }
        }
    if (seg_min < 0) return 0;
    tp_min *= 0.9999; // KM:562
    p.x = xo + tp_min * (p.x - xo);
    p.y = yo + tp_min * (p.y - yo);
    p.li = (p.x - m.x0) / m.dhx;
    p.lj = (p.y - m.y0) / m.dhy;
    p.dt = dt0 * (1 - tp_min);
    if (p.li < 0 && p.li > -SF_FLT_EPS) p.li = 0; // KM:574-577
    if (p.lj < 0 && p.lj > -SF_FLT_EPS) p.lj = 0;
    const int2 kd = m.seg_kind[seg_min];
    const bool alive = kd.x != 0 && kd.y == 0; // performSurfaceInteraction KM:586-587, SINK KM:593-594
    if (kd.x == 2) { // SurfaceImpactSpecular without a species change (SurfaceInteraction.java:104-149): vel += n * (|vel_xy| * sqrt 2), before the hit is recorded
        const double4 sg = m.seg_xy[seg_min];
        double dx = sg.z - sg.x, dy = sg.w - sg.y; // LinearSegment.normal, LinearSegment.java:26-44
        const double len = sqrt(dx * dx + dy * dy);
        dx /= len;
        dy /= len;
        const double n0 = -dy, n1 = dx;
        const double mag = sqrt(p.u * p.u + p.v * p.v) * 1.4142135623730951; // Vec.mag2 * Constants.SQRT2 (= Math.sqrt(2))
        p.u += n0 * mag;
        p.v += n1 * mag;
    }
    if (m.hits.n) {
        const unsigned long long h = atomicAdd(m.hits.n, 1ULL);
        if (h < m.hits.cap) {
            m.hits.seg[h] = seg_min; m.hits.mesh[h] = m.id; m.hits.t[h] = tsurf_min;
            m.hits.u[h] = p.u; m.hits.v[h] = p.v; m.hits.w[h] = p.w; m.hits.mpw[h] = p.mpw;
            m.hits.alive[h] = alive ? 1 : 0;
        }
    }
    return alive ? 1 : 2;
}

// MESH:1476-1483 on a uniform mesh
__device__ __forceinline__ bool sf_contains_pos(const MeshDev &m, double x, double y, double &li, double &lj)
{
    li = (x - m.x0) / m.dhx;
    lj = (y - m.y0) / m.dhy;
    return !(li < -SF_FLT_EPS || lj < -SF_FLT_EPS || li > (m.ni - 1 + SF_FLT_EPS) || lj > (m.nj - 1 + SF_FLT_EPS));
}

// ParticleMover.run for one particle, KM:317-420, with ProcessBoundary's domain part, KM:605-749.
// `exact_lc` is cleared when the stored lc stops being XtoL(pos) (boundary clamp, stale periodic lc).
template <class FieldGather>
__device__ __forceinline__ int sf_move(const MeshDev &m, const MeshDev *__restrict__ meshes, double qm, double charge,
                                       double dt, bool transfer, PState &p, MoveAux &aux, bool &exact_lc,
                                       const FieldGather &fg)
{
    aux.bounces = 0;
    aux.xfer_mask = 0;
    if (p.mpw <= 0) return SF_REMOVED; // KM:322
    if (!transfer) {                   // KM:332-354
        p.dt += dt;
        sf_kick(m, qm, p.dt, p, fg);
    }
    const int ni = m.ni, nj = m.nj;
    int bounces = 0;
    while (p.dt > 0 && bounces++ < SF_MAX_BOUNCES) { // KM:360
        const double xo = p.x, yo = p.y, lio = p.li, ljo = p.lj;
        p.x += p.u * p.dt; // KM:369-370
        p.y += p.v * p.dt;
        if (m.domain == SFGPU_RZ) { // rotateToRZ, KM:424-442
            double A = p.w * p.dt, B = p.x, R = sqrt(A * A + B * B);
            double c = B / R, s = A / R;
            p.z -= asin(s);
            double v1 = p.u, v2 = p.w;
            p.x = R;
            p.u = c * v1 + s * v2;
            p.w = -s * v1 + c * v2;
        } else if (m.domain == SFGPU_ZR) { // rotateToZR, KM:444-462
            double A = p.w * p.dt, B = p.y, R = sqrt(A * A + B * B);
            double c = B / R, s = A / R;
            p.z += acos(c);
            double v1 = p.v, v2 = p.w;
            p.y = R;
            p.v = c * v1 + s * v2;
            p.w = -s * v1 + c * v2;
        } else {
            p.z += p.w * p.dt; // KM:380
        }
        p.li = (p.x - m.x0) / m.dhx; // UM:158-159: true divisions
        p.lj = (p.y - m.y0) / m.dhy;
        exact_lc = true;

        // ---- ProcessBoundary ----
        const bool near_segments = m.any_seg && sf_bbox_has_segments(m, p.li, p.lj, lio, ljo);
        if (near_segments && !m.seg_offs) {
            aux.xo = xo; aux.yo = yo; aux.lio = lio; aux.ljo = ljo;
            aux.bounces = bounces;
            return SF_SLOW; // pre-ProcessBoundary state, dt still holds dt0
        }
        const double dt0 = p.dt;
        p.dt = 0; // KM:475-476
        if (near_segments) {
            const int hit = sf_process_segments(&m, dt0, xo, yo, lio, ljo, &p);
            if (hit == 2) return SF_ABSORBED;
            if (hit == 1) exact_lc = (p.li == (p.x - m.x0) / m.dhx) && (p.lj == (p.y - m.y0) / m.dhy); // (the tiny-negative clamp of KM:574-577)
        }
        if (p.li < 0 || p.lj < 0 || p.li >= ni - 1 || p.lj >= nj - 1) { // KM:606
            const double xs = p.x, ys = p.y, lis = p.li, ljs = p.lj;
            double t_right = 99, t_top = 99, t_left = 99, t_bottom = 99;
            if (p.li >= ni - 1) t_right = (ni - 1.0 - lio) / (p.li - lio);
            if (p.lj >= nj - 1) t_top = (nj - 1.0 - ljo) / (p.lj - ljo);
            if (p.li < 0) t_left = lio / (lio - p.li);
            if (p.lj < 0) t_bottom = ljo / (ljo - p.lj);
            int face = SFGPU_FACE_RIGHT;
            double t = t_right;
            if (t_top < t) { face = SFGPU_FACE_TOP; t = t_top; }
            if (t_left < t) { face = SFGPU_FACE_LEFT; t = t_left; }
            if (t_bottom < t) { face = SFGPU_FACE_BOTTOM; t = t_bottom; }
            p.li = lio + t * (p.li - lio); // KM:644-645
            p.lj = ljo + t * (p.lj - ljo);
            if (p.li < 0) p.li = 0; else if (p.li > ni - 1) p.li = ni - 1;
            if (p.lj < 0) p.lj = 0; else if (p.lj > nj - 1) p.lj = nj - 1;
            p.x = m.x0 + p.li * m.dhx; // mesh.pos(lc), UM:139-145
            p.y = m.y0 + p.lj * m.dhy;
            p.dt = dt0 * (1 - t); // KM:663
            exact_lc = false;
            int i = sf_j2i(p.li), j = sf_j2i(p.lj);
            if (face == SFGPU_FACE_TOP) j++;
            if (face == SFGPU_FACE_RIGHT) i++;
            if (i < 0) i = 0;
            if (j < 0) j = 0;
            if (i >= ni - 1) i = ni - 1;
            if (j >= nj - 1) j = nj - 1;
            const int type = (face == SFGPU_FACE_LEFT || face == SFGPU_FACE_RIGHT) ? m.bc[face][j] : m.bc[face][i];
            switch (type) {
            case SFGPU_BC_SYMMETRY: sf_mirror(p, face); break; // KM:692-696
            case SFGPU_BC_PERIODIC:                             // KM:697-707
                if (face == SFGPU_FACE_LEFT) p.x += m.lenx;
                else if (face == SFGPU_FACE_RIGHT) p.x -= m.lenx;
                else if (face == SFGPU_FACE_BOTTOM) p.y += m.leny;
                else p.y -= m.leny;
                break;
            case SFGPU_BC_MESH: { // KM:708-722
                const int index = (face == SFGPU_FACE_LEFT || face == SFGPU_FACE_RIGHT) ? sf_j2i(p.lj) : sf_j2i(p.li);
                int mask = 0;
                for (int k = 0; k < 2; k++) {
                    const int nb = m.nbr[face] ? m.nbr[face][2 * index + k] : -1;
                    double nli, nlj;
                    if (nb >= 0 && sf_contains_pos(meshes[nb], p.x, p.y, nli, nlj)) {
                        mask |= 1 << k;
                        aux.xfer_mesh[k] = nb;
                        aux.xfer_li[k] = nli;
                        aux.xfer_lj[k] = nlj;
                    }
                }
                aux.xfer_mask = mask;
                aux.bounces = bounces;
                return SF_TRANSFER;
            }
            case SFGPU_BC_CIRCUIT: // KM:723-742: ions die, electrons need the global wall charge
                if (charge >= 0) return SF_DEAD;
                p.x = xs; p.y = ys; p.li = lis; p.lj = ljs; p.dt = dt0;
                exact_lc = true;
                aux.xo = xo; aux.yo = yo; aux.lio = lio; aux.ljo = ljo;
                aux.bounces = bounces;
                return SF_SLOW;
            default: // OPEN and everything else, KM:690-691, :744-745
                return SF_DEAD;
            }
        }
    }
    aux.bounces = bounces > SF_MAX_BOUNCES ? SF_MAX_BOUNCES : bounces;
    return SF_ALIVE;
}

// bilinear deposit weights of F2D:244-295 incl. the Ruyten correction; false when scatter() returns early
struct DepW {
    int i, j;
    double w00, w10, w11, w01; // nodes (i,j) (i+1,j) (i+1,j+1) (i,j+1)
};

__device__ __forceinline__ bool sf_deposit_weights(const MeshDev &m, double fi, double fj, DepW &d)
{
    const int i = sf_j2i(fi), j = sf_j2i(fj);
    d.i = i;
    d.j = j;
    if (i < 0 || j < 0 || i >= m.ni - 1 || j >= m.nj - 1) return false;
    double di = fi - i, dj = fj - j;
    if (m.domain == SFGPU_RZ) {
        double rp = m.x0 + (i + 1) * m.dhx, rm = m.x0 + i * m.dhx, r = m.x0 + fi * m.dhx;
        di = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    } else if (m.domain == SFGPU_ZR) {
        double rp = m.y0 + (j + 1) * m.dhy, rm = m.y0 + j * m.dhy, r = m.y0 + fj * m.dhy;
        dj = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    }
    d.w00 = (1 - di) * (1 - dj);
    d.w10 = di * (1 - dj);
    d.w11 = di * dj;
    d.w01 = (1 - di) * dj;
    return true;
}

// the seven bilinear deposit values of one particle: KM:184-187, KM:1584-1590
__device__ __forceinline__ void sf_deposit_values(const PState &p, double val[7])
{
    val[0] = p.mpw;
    val[1] = p.mpw * p.u;
    val[2] = p.mpw * p.v;
    val[3] = p.mpw * p.w;
    val[4] = val[1] * p.u;
    val[5] = val[2] * p.v;
    val[6] = val[3] * p.w;
}
