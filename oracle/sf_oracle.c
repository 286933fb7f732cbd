/*
 * sf_oracle.c -- CPU ORACLE (test infrastructure, see sf_oracle.h).  PARITY UNPINNED by the
 * reference (it has no tests); pinned by tests/test_oracle_kat.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread (oracle/Makefile).
 * Every function cites the Java lines it restates.
 */
#include "sf_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FLT_EPS 1e-7 /* Constants.java:26 */

/* Java (int)double: toward zero, saturating, NaN -> 0 (JLS 5.1.3) */
static inline int j2i(double d)
{
    if (d != d) return 0;
    if (d >= 2147483647.0) return INT_MAX;
    if (d <= -2147483648.0) return INT_MIN;
    return (int)d;
}

/* F2D:371-390 */
double sfo_gather_safe(const double *d, int ni, int nj, double fi, double fj)
{
    int i = j2i(fi), j = j2i(fj);
    double di = fi - i, dj = fj - j, v;
    if (i < 0) { i = 0; di = 0; }
    if (j < 0) { j = 0; dj = 0; }
    if (i >= ni - 1) { i = ni - 1; di = 0; }
    if (j >= nj - 1) { j = nj - 1; dj = 0; }
    v = (1 - di) * (1 - dj) * d[(int64_t)i * nj + j];
    if (di > 0) v += di * (1 - dj) * d[(int64_t)(i + 1) * nj + j];
    if (di > 0 && dj > 0) v += di * dj * d[(int64_t)(i + 1) * nj + j + 1];
    if (dj > 0) v += (1 - di) * dj * d[(int64_t)i * nj + j + 1];
    return v;
}

/* F2D:300-350: gather(), falling back to gather_safe() when Java would have thrown
 * IndexOutOfBoundsException, i.e. when any of the four node indices is outside the array. */
double sfo_gather(const double *d, int ni, int nj, double fi, double fj)
{
    int i = j2i(fi), j = j2i(fj);
    if (i < 0 || j < 0 || i >= ni - 1 || j >= nj - 1) return sfo_gather_safe(d, ni, nj, fi, fj);
    double di = fi - i, dj = fj - j, v;
    v = (1 - di) * (1 - dj) * d[(int64_t)i * nj + j];
    v += di * (1 - dj) * d[(int64_t)(i + 1) * nj + j];
    v += di * dj * d[(int64_t)(i + 1) * nj + j + 1];
    v += (1 - di) * dj * d[(int64_t)i * nj + j + 1];
    return v;
}

/* UM:154-161 */
void sfo_xtol(const sfo_mesh *m, double x, double y, double *li, double *lj)
{
    *li = (x - m->x0[0]) / m->dh[0];
    *lj = (y - m->x0[1]) / m->dh[1];
}

/* MESH:1476-1483 */
int sfo_contains_pos(const sfo_mesh *m, double x, double y)
{
    double li, lj;
    sfo_xtol(m, x, y, &li, &lj);
    if (li < -FLT_EPS || lj < -FLT_EPS || li > (m->ni - 1 + FLT_EPS) || lj > (m->nj - 1 + FLT_EPS))
        return 0;
    return 1;
}

/* F2D:244-295; mesh.R() = pos1 (RZ) / pos2 (ZR), MESH:824-831 with UM:139-145 */
void sfo_scatter(double *d, const sfo_mesh *m, double fi, double fj, double val)
{
    int ni = m->ni, nj = m->nj;
    int i = j2i(fi), j = j2i(fj);
    double di = fi - i, dj = fj - j;
    if (i < 0 || j < 0 || i >= ni - 1 || j >= nj - 1) return;
    if (m->domain_type == SFO_RZ) {
        double rp = m->x0[0] + (i + 1) * m->dh[0];
        double rm = m->x0[0] + i * m->dh[0];
        double r = m->x0[0] + fi * m->dh[0];
        di = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    } else if (m->domain_type == SFO_ZR) {
        double rp = m->x0[1] + (j + 1) * m->dh[1];
        double rm = m->x0[1] + j * m->dh[1];
        double r = m->x0[1] + fj * m->dh[1];
        di = fi - i;
        dj = 1 - (0.5 * (rp - r) * (2 * rp + 3 * rm - r) / (rp * rp - rm * rm));
    }
    d[(int64_t)i * nj + j] += (1 - di) * (1 - dj) * val;
    d[(int64_t)(i + 1) * nj + j] += di * (1 - dj) * val;
    d[(int64_t)(i + 1) * nj + j + 1] += di * dj * val;
    d[(int64_t)i * nj + j + 1] += (1 - di) * dj * val;
}

/* Vec.java:318-325 */
static void cross3(const double a[3], const double b[3], double r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = -a[0] * b[2] + a[2] * b[0];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

/* KM:847-893 */
void sfo_boris(double qm, double dtp, const double E[3], const double B[3], double vel[3])
{
    double t[3], s[3], vm[3], vp[3], vpl[3], c[3], tm2;
    int k;
    for (k = 0; k < 3; k++) t[k] = qm * B[k] * 0.5 * dtp;
    tm2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    for (k = 0; k < 3; k++) s[k] = 2 * t[k] / (1 + tm2);
    for (k = 0; k < 3; k++) vm[k] = vel[k] + qm * E[k] * 0.5 * dtp;
    cross3(vm, t, c);
    for (k = 0; k < 3; k++) vp[k] = vm[k] + c[k];
    cross3(vp, s, c);
    for (k = 0; k < 3; k++) vpl[k] = vm[k] + c[k];
    for (k = 0; k < 3; k++) vel[k] = vpl[k] + qm * E[k] * 0.5 * dtp;
}

/* Vec.java:406-418 with dot (:151-159), mult (:71-77), subtract (:43-51), add (:57-65) */
void sfo_mirror(double vel[3], const double n[3])
{
    double tm = 0, t[3], nn[3];
    int k;
    for (k = 0; k < 3; k++) tm += vel[k] * n[k];
    for (k = 0; k < 3; k++) t[k] = n[k] * tm;
    for (k = 0; k < 3; k++) nn[k] = vel[k] - t[k];
    for (k = 0; k < 3; k++) t[k] = t[k] * -1;
    for (k = 0; k < 3; k++) vel[k] = t[k] + nn[k];
}

/* velocity kick shared by KM:336-353 and KM:782-794 */
static void kick(const sfo_mesh *m, double qm, double dtp, double li, double lj, double vel[3])
{
    double ef[3] = {0, 0, 0}, bf[3] = {0, 0, 0};
    ef[0] = sfo_gather(m->efi, m->ni, m->nj, li, lj);
    ef[1] = sfo_gather(m->efj, m->ni, m->nj, li, lj);
    if (m->bfi) bf[0] = sfo_gather(m->bfi, m->ni, m->nj, li, lj);
    if (m->bfj) bf[1] = sfo_gather(m->bfj, m->ni, m->nj, li, lj);
    if (bf[0] == 0 && bf[1] == 0) {
        vel[0] += qm * ef[0] * dtp;
        vel[1] += qm * ef[1] * dtp;
    } else {
        sfo_boris(qm, dtp, ef, bf, vel);
    }
}

/* does the node bounding box of the substep hold a DIRICHLET/SINK segment? KM:482-518 */
static int bbox_has_segments(const sfo_mesh *m, double li, double lj, double lio, double ljo)
{
    if (!m->has_seg) return 0;
    /* Math.min/max: NaN if either is NaN; (int)NaN = 0 */
    double mn0 = (li != li || lio != lio) ? NAN : (li < lio ? li : lio);
    double mn1 = (lj != lj || ljo != ljo) ? NAN : (lj < ljo ? lj : ljo);
    double mx0 = (li != li || lio != lio) ? NAN : (li > lio ? li : lio);
    double mx1 = (lj != lj || ljo != ljo) ? NAN : (lj > ljo ? lj : ljo);
    int i_min = j2i(mn0), i_max = j2i(mx0), j_min = j2i(mn1), j_max = j2i(mx1);
    if (i_min < 0) i_min = 0;
    if (j_min < 0) j_min = 0;
    if (i_max >= m->ni) i_max = m->ni - 1;
    if (j_max >= m->nj) j_max = m->nj - 1;
    for (int i = i_min; i <= i_max; i++)
        for (int j = j_min; j <= j_max; j++)
            if (m->has_seg[(int64_t)i * m->nj + j]) return 1;
    return 0;
}

/* LinearSegment.InfiniteLineIntersect + intersect, LinearSegment.java:113-179 */
void sfo_segment_intersect(const sfo_segment *s, const double p3[2], const double p4[2], double t[2])
{
    const double x1 = s->x1, x2 = s->x2, x3 = p3[0], x4 = p4[0];
    const double y1 = s->y1, y2 = s->y2, y3 = p3[1], y4 = p4[1];
    const double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    t[0] = -1;
    t[1] = -1;
    if (den == 0) return;
    const double xp0 = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / den;
    const double xp1 = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / den;
    double t0, t1;
    if (fabs(x2 - x1) > 1e-6) t0 = (xp0 - x1) / (x2 - x1);
    else t0 = (xp1 - y1) / (y2 - y1);
    if (t0 < -FLT_EPS || t0 > (1 + FLT_EPS)) return;
    if (fabs(x4 - x3) > 1e-6) t1 = (xp0 - x3) / (x4 - x3);
    else t1 = (xp1 - y3) / (y4 - y3);
    if (t1 < -FLT_EPS || t1 > (1 + FLT_EPS)) return;
    if (t0 < 0) t0 = 0;
    if (t1 < 0) t1 = 0;
    if (t0 > 1) t0 = 1;
    if (t1 > 1) t1 = 1;
    t[0] = t0;
    t[1] = t1;
}

/* segment part of ProcessBoundary, KM:504-603, for LinearSegments with a deterministic surface outcome.
 * Returns 0: no hit, 1: hit and alive (pos, lc, *dtp updated), 2: hit and removed. */
static int process_segments(const sfo_mesh *m, double dt0, const double old[2], double lio, double ljo, double pos[3], double vel[3],
                            double mpw, double *li, double *lj, double *dtp)
{
    /* the node bounding box of the sub-step, KM:482-502 */
    double mn0 = (*li != *li || lio != lio) ? NAN : (*li < lio ? *li : lio);
    double mn1 = (*lj != *lj || ljo != ljo) ? NAN : (*lj < ljo ? *lj : ljo);
    double mx0 = (*li != *li || lio != lio) ? NAN : (*li > lio ? *li : lio);
    double mx1 = (*lj != *lj || ljo != ljo) ? NAN : (*lj > ljo ? *lj : ljo);
    int i_min = j2i(mn0), i_max = j2i(mx0), j_min = j2i(mn1), j_max = j2i(mx1);
    if (i_min < 0) i_min = 0;
    if (j_min < 0) j_min = 0;
    if (i_max >= m->ni) i_max = m->ni - 1;
    if (j_max >= m->nj) j_max = m->nj - 1;
    double tp_min = 2.0, tsurf_min = 0;
    int seg_min = -1;
    for (int i = i_min; i <= i_max; i++)
        for (int j = j_min; j <= j_max; j++) {
            const int64_t node = (int64_t)i * m->nj + j;
            for (int k = m->seg_offs[node]; k < m->seg_offs[node + 1]; k++) { /* (a segment met twice gives the same t: the set of KM:505 is not needed) */
                const sfo_segment *seg = &m->segs[m->seg_ids[k]];
                double t[2];
                sfo_segment_intersect(seg, old, pos, t);
                const double t_part = t[1];
                if (t_part > 0) { /* KM:535 */
                    double dx = seg->x2 - seg->x1, dy = seg->y2 - seg->y1; /* LinearSegment.java:26-44 */
                    const double len = sqrt(dx * dx + dy * dy);
                    dx /= len;
                    dy /= len;
                    const double acos_ = (-dy * vel[0] + dx * vel[1]) / sqrt(vel[0] * vel[0] + vel[1] * vel[1]);
                    if (t_part < FLT_EPS && acos_ > 0) continue; /* KM:541-544 */
                    if (t_part < tp_min) {
                        tp_min = t_part;
                        tsurf_min = t[0];
                        seg_min = m->seg_ids[k];
                    }
                }
            }
        }
    if (seg_min < 0) return 0;
    tp_min *= 0.9999; /* KM:562 */
    pos[0] = old[0] + tp_min * (pos[0] - old[0]);
    pos[1] = old[1] + tp_min * (pos[1] - old[1]);
    sfo_xtol(m, pos[0], pos[1], li, lj);
    *dtp = dt0 * (1 - tp_min);
    if (*li < 0 && *li > -FLT_EPS) *li = 0; /* KM:574-577 */
    if (*lj < 0 && *lj > -FLT_EPS) *lj = 0;
    const sfo_segment *seg = &m->segs[seg_min];
    int alive = seg->kind != 0; /* performSurfaceInteraction, KM:586-587 (Material.java:279-300): nothing listed / ABSORB remove, NONE keeps */
    if (seg->kind == 2) { /* SurfaceImpactSpecular without a species change, SurfaceInteraction.java:104-149: the velocity the hit records carry is the new one */
        double dx = seg->x2 - seg->x1, dy = seg->y2 - seg->y1; /* LinearSegment.normal, LinearSegment.java:26-44 */
        const double len = sqrt(dx * dx + dy * dy);
        dx /= len;
        dy /= len;
        const double n0 = -dy, n1 = dx;
        const double mag = sqrt(vel[0] * vel[0] + vel[1] * vel[1]) * sqrt(2.0); /* Vec.mag2 (Vec.java:269-272) * Constants.SQRT2 */
        vel[0] += n0 * mag;
        vel[1] += n1 * mag;
    }
    if (seg->sink) alive = 0;   /* KM:593-594 */
    if (m->hits) {
        const int64_t h = __sync_fetch_and_add(&m->hits->n, 1);
        if (h < m->hits->cap) {
            m->hits->seg[h] = seg_min; m->hits->t[h] = tsurf_min;
            m->hits->u[h] = vel[0]; m->hits->v[h] = vel[1]; m->hits->w[h] = vel[2];
            m->hits->mpw[h] = mpw; m->hits->alive[h] = (int8_t)alive;
        }
    }
    return alive ? 1 : 2;
}

/* ParticleMover.run, KM:298-422 with ProcessBoundary's domain-exit part, KM:605-749 */
void sfo_move(const sfo_mesh *meshes, int mesh_id, double qm, double charge, double dt,
              int particle_transfer, sfo_particles *p, int64_t first, int64_t count,
              sfo_move_out *out, double sums5[5])
{
    const sfo_mesh *m = &meshes[mesh_id];
    const int ni = m->ni, nj = m->nj;
    const double xd0 = m->x0[0] + (ni - 1) * m->dh[0]; /* UM:131-135 */
    const double xd1 = m->x0[1] + (nj - 1) * m->dh[1];
    double N_sum = 0, P0 = 0, P1 = 0, P2 = 0, E_sum = 0;

    for (int64_t q = first; q < first + count; q++) {
        int8_t st = SFO_ALIVE;
        if (out->xfer_mask) out->xfer_mask[q] = 0;
        if (out->bounces) out->bounces[q] = 0;
        if (p->mpw[q] <= 0) { /* KM:322 */
            out->status[q] = SFO_REMOVED;
            continue;
        }
        double pos[3] = {p->x[q], p->y[q], p->z[q]};
        double vel[3] = {p->u[q], p->v[q], p->w[q]};
        double li = p->li[q], lj = p->lj[q], dtp = p->dt[q];

        if (!particle_transfer) { /* KM:332-354 */
            dtp += dt;
            kick(m, qm, dtp, li, lj, vel);
        }

        int bounces = 0, alive = 1;
        while (dtp > 0 && bounces++ < 10) { /* KM:360 */
            double xo = pos[0], yo = pos[1], lio = li, ljo = lj;
            pos[0] += vel[0] * dtp; /* KM:369-370 */
            pos[1] += vel[1] * dtp;
            if (m->domain_type == SFO_RZ) { /* KM:424-442 */
                double A = vel[2] * dtp, B = pos[0], R = sqrt(A * A + B * B);
                double c = B / R, s = A / R;
                pos[2] -= asin(s);
                double v1 = vel[0], v2 = vel[2];
                pos[0] = R;
                vel[0] = c * v1 + s * v2;
                vel[2] = -s * v1 + c * v2;
            } else if (m->domain_type == SFO_ZR) { /* KM:444-462 */
                double A = vel[2] * dtp, B = pos[1], R = sqrt(A * A + B * B);
                double c = B / R, s = A / R;
                pos[2] += acos(c);
                double v1 = vel[1], v2 = vel[2];
                pos[1] = R;
                vel[1] = c * v1 + s * v2;
                vel[2] = -s * v1 + c * v2;
            } else {
                pos[2] += vel[2] * dtp; /* KM:380 */
            }
            sfo_xtol(m, pos[0], pos[1], &li, &lj); /* KM:384 */

            /* ---- ProcessBoundary, KM:471-750 ---- */
            const int near_segments = bbox_has_segments(m, li, lj, lio, ljo);
            if (near_segments && !m->segs) {
                /* segment intersection + surface interaction stay in Java; hand the particle
                 * over in its pre-ProcessBoundary state */
                st = SFO_SLOW;
                if (out->old_x) { out->old_x[q] = xo; out->old_y[q] = yo; out->old_li[q] = lio; out->old_lj[q] = ljo; }
                alive = 0;
                break;
            }
            const double dt0 = dtp;
            dtp = 0; /* KM:475-476 */
            if (near_segments) {
                const double old[2] = {xo, yo};
                if (process_segments(m, dt0, old, lio, ljo, pos, vel, p->mpw[q], &li, &lj, &dtp) == 2) {
                    st = SFO_ABSORBED;
                    alive = 0;
                    break;
                }
            }
            const double xs = pos[0], ys = pos[1], lis = li, ljs = lj;
            if (li < 0 || lj < 0 || li >= ni - 1 || lj >= nj - 1) { /* KM:606 */
                double t_right = 99, t_top = 99, t_left = 99, t_bottom = 99;
                if (li >= ni - 1) t_right = (ni - 1.0 - lio) / (li - lio);
                if (lj >= nj - 1) t_top = (nj - 1.0 - ljo) / (lj - ljo);
                if (li < 0) t_left = lio / (lio - li);
                if (lj < 0) t_bottom = ljo / (ljo - lj);
                int face = SFO_RIGHT;
                double t = t_right;
                if (t_top < t) { face = SFO_TOP; t = t_top; }
                if (t_left < t) { face = SFO_LEFT; t = t_left; }
                if (t_bottom < t) { face = SFO_BOTTOM; t = t_bottom; }
                li = lio + t * (li - lio); /* KM:644-645 */
                lj = ljo + t * (lj - ljo);
                if (li < 0) li = 0; else if (li > ni - 1) li = ni - 1;
                if (lj < 0) lj = 0; else if (lj > nj - 1) lj = nj - 1;
                pos[0] = m->x0[0] + li * m->dh[0]; /* mesh.pos(lc), UM:139-145 */
                pos[1] = m->x0[1] + lj * m->dh[1];
                dtp = dt0 * (1 - t); /* KM:663 */
                int i = j2i(li), j = j2i(lj);
                if (face == SFO_TOP) j++;
                if (face == SFO_RIGHT) i++;
                if (i < 0) i = 0;
                if (j < 0) j = 0;
                if (i >= ni - 1) i = ni - 1;
                if (j >= nj - 1) j = nj - 1;
                int type = (face == SFO_LEFT || face == SFO_RIGHT) ? m->bc[face][j] : m->bc[face][i];
                switch (type) {
                case SFO_OPEN: alive = 0; st = SFO_DEAD; break;
                case SFO_SYMMETRY: { /* KM:692-696, UM:174-187 */
                    double n[3] = {0, 0, 0};
                    if (face == SFO_LEFT) n[0] = 1;
                    else if (face == SFO_RIGHT) n[0] = -1;
                    else if (face == SFO_BOTTOM) n[1] = 1;
                    else n[1] = -1;
                    sfo_mirror(vel, n);
                    break;
                }
                case SFO_PERIODIC: /* KM:697-707 */
                    if (face == SFO_LEFT) pos[0] += (xd0 - m->x0[0]);
                    else if (face == SFO_RIGHT) pos[0] -= (xd0 - m->x0[0]);
                    else if (face == SFO_BOTTOM) pos[1] += (xd1 - m->x0[1]);
                    else pos[1] -= (xd1 - m->x0[1]);
                    break;
                case SFO_MESH: { /* KM:708-722 */
                    int index = (face == SFO_LEFT || face == SFO_RIGHT) ? j2i(lj) : j2i(li);
                    int mask = 0;
                    for (int k = 0; k < 2; k++) {
                        int nb = m->nbr[face] ? m->nbr[face][2 * index + k] : -1;
                        if (nb >= 0 && sfo_contains_pos(&meshes[nb], pos[0], pos[1])) {
                            sfo_xtol(&meshes[nb], pos[0], pos[1], &li, &lj);
                            mask |= 1 << k;
                            if (out->xfer_mesh) {
                                out->xfer_mesh[2 * q + k] = nb;
                                out->xfer_li[2 * q + k] = li;
                                out->xfer_lj[2 * q + k] = lj;
                            }
                        }
                    }
                    if (out->xfer_mask) out->xfer_mask[q] = mask;
                    alive = 0;
                    st = SFO_TRANSFER;
                    break;
                }
                case SFO_CIRCUIT: /* KM:723-742: ions die; electrons depend on the global wall charge */
                    alive = 0;
                    st = (charge >= 0) ? SFO_DEAD : SFO_SLOW;
                    if (st == SFO_SLOW) { /* hand over in the pre-ProcessBoundary state */
                        pos[0] = xs; pos[1] = ys; li = lis; lj = ljs; dtp = dt0;
                        if (out->old_x) { out->old_x[q] = xo; out->old_y[q] = yo; out->old_li[q] = lio; out->old_lj[q] = ljo; }
                    }
                    break;
                default: alive = 0; st = SFO_DEAD; break;
                }
            }
            if (!alive) break; /* KM:393-396 */
        }
        if (alive) { /* KM:406-413 */
            N_sum += p->mpw[q];
            P0 += p->mpw[q] * vel[0];
            P1 += p->mpw[q] * vel[1];
            P2 += p->mpw[q] * vel[2];
            E_sum += p->mpw[q] * sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
        }
        p->x[q] = pos[0]; p->y[q] = pos[1]; p->z[q] = pos[2];
        p->u[q] = vel[0]; p->v[q] = vel[1]; p->w[q] = vel[2];
        p->li[q] = li; p->lj[q] = lj; p->dt[q] = dtp;
        out->status[q] = st;
        if (out->bounces) out->bounces[q] = bounces > 10 ? 10 : bounces;
    }
    sums5[0] += N_sum; sums5[1] += P0; sums5[2] += P1; sums5[3] += P2; sums5[4] += E_sum;
}

typedef struct {
    const sfo_mesh *meshes; int mesh_id; double qm, charge, dt; int transfer;
    sfo_particles *p; int64_t first, count; sfo_move_out *out; double sums[5];
} mt_arg;

static void *mt_run(void *a_)
{
    mt_arg *a = (mt_arg *)a_;
    sfo_move(a->meshes, a->mesh_id, a->qm, a->charge, a->dt, a->transfer, a->p, a->first, a->count, a->out, a->sums);
    return NULL;
}

/* KM:200-261: one ParticleMover thread per block; sums reduced in block order */
void sfo_move_mt(const sfo_mesh *meshes, int mesh_id, double qm, double charge, double dt,
                 int particle_transfer, sfo_particles *p, sfo_move_out *out, double sums5[5], int threads)
{
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    mt_arg *args = (mt_arg *)calloc(threads, sizeof(mt_arg));
    int64_t per = (p->n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        int64_t f = per * t, c = p->n - f;
        if (c > per) c = per;
        if (c < 0) c = 0;
        mt_arg a = {meshes, mesh_id, qm, charge, dt, particle_transfer, p, f, c, out, {0, 0, 0, 0, 0}};
        args[t] = a;
        pthread_create(&th[t], NULL, mt_run, &args[t]);
    }
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        for (int k = 0; k < 5; k++) sums5[k] += args[t].sums[k];
    }
    free(th);
    free(args);
}

/* KM:180-188 */
void sfo_deposit(const sfo_mesh *m, const sfo_particles *p, double *den, double *u, double *v, double *w)
{
    for (int64_t q = 0; q < p->n; q++) {
        sfo_scatter(den, m, p->li[q], p->lj[q], p->mpw[q]);
        sfo_scatter(u, m, p->li[q], p->lj[q], p->u[q] * p->mpw[q]);
        sfo_scatter(v, m, p->li[q], p->lj[q], p->v[q] * p->mpw[q]);
        sfo_scatter(w, m, p->li[q], p->lj[q], p->w[q] * p->mpw[q]);
    }
}

/* F2D:403-414 */
void sfo_divide_by_field(double *d, const double *by, int64_t n)
{
    for (int64_t k = 0; k < n; k++) {
        if (by[k] != 0) d[k] /= by[k];
        else d[k] = 0;
    }
}

/* F2D:418-431 */
void sfo_scale_by_vol(double *d, const double *node_vol, int64_t n)
{
    for (int64_t k = 0; k < n; k++) d[k] /= node_vol[k];
}

/* KM:1580-1594 */
void sfo_sample(const sfo_mesh *m, const sfo_particles *p, double *count_sum, double *u_sum, double *v_sum,
                double *w_sum, double *uu_sum, double *vv_sum, double *ww_sum, double *mpc_sum)
{
    for (int64_t q = 0; q < p->n; q++) {
        double li = p->li[q], lj = p->lj[q], mpw = p->mpw[q];
        sfo_scatter(u_sum, m, li, lj, mpw * p->u[q]);
        sfo_scatter(v_sum, m, li, lj, mpw * p->v[q]);
        sfo_scatter(w_sum, m, li, lj, mpw * p->w[q]);
        sfo_scatter(uu_sum, m, li, lj, mpw * p->u[q] * p->u[q]);
        sfo_scatter(vv_sum, m, li, lj, mpw * p->v[q] * p->v[q]);
        sfo_scatter(ww_sum, m, li, lj, mpw * p->w[q] * p->w[q]);
        sfo_scatter(count_sum, m, li, lj, mpw);
        int ci = j2i(li), cj = j2i(lj); /* Java would throw outside the array; skip instead */
        if (ci >= 0 && cj >= 0 && ci < m->ni && cj < m->nj) mpc_sum[(int64_t)ci * m->nj + cj] += 1;
    }
}

/* KM:759-802 */
void sfo_add_particles(const sfo_mesh *m, double qm, double dt, int compute_lc, sfo_particles *p,
                       int64_t first, int64_t count)
{
    for (int64_t q = first; q < first + count; q++) {
        if (compute_lc) {
            sfo_xtol(m, p->x[q], p->y[q], &p->li[q], &p->lj[q]);
            if (p->li[q] >= m->ni) p->li[q] = m->ni - 1;
            if (p->lj[q] >= m->nj) p->lj[q] = m->nj - 1;
        }
        double vel[3] = {p->u[q], p->v[q], p->w[q]};
        kick(m, qm, -0.5 * dt, p->li[q], p->lj[q], vel);
        p->u[q] = vel[0]; p->v[q] = vel[1]; p->w[q] = vel[2];
        p->dt[q] = 0;
    }
}

/* ---------------------------------------------------------------------------------------------
 * particle injection by UniformSource / ColdBeamSource on a Boundary of linear segments (XY domains)
 * ------------------------------------------------------------------------------------------- */

/* java.util.Random: 48-bit LCG, next(bits) (the JDK's documented algorithm; Starfish.rnd() = random.nextDouble(),
 * Starfish.java:244-246).  `state` is the scrambled internal seed ((seed ^ 0x5DEECE66D) & (2^48 - 1) after setSeed). */
static int32_t java_next(uint64_t *state, int bits)
{
    *state = (*state * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)((int64_t)*state >> (48 - bits)); /* (int)(seed >>> (48 - bits)) */
}
uint64_t sfo_java_seed(int64_t seed) { return ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
int32_t sfo_java_next_int(uint64_t *state) { return java_next(state, 32); }
double sfo_java_next_double(uint64_t *state)
{
    const int64_t hi = (int64_t)java_next(state, 26), lo = (int64_t)java_next(state, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53; /* DOUBLE_UNIT */
}

/* Vec.binarySearch, Vec.java:529-546 */
static int vec_binary_search(const double *vec, int n, double val)
{
    if (val < vec[0]) return -1;
    if (val > vec[n - 1]) return n;
    int i1 = 0, i2 = n;
    for (;;) {
        const int i_mid = (int)(0.5 * (i1 + i2));
        if (val < vec[i_mid]) i2 = i_mid;
        else if (val > vec[i_mid]) i1 = i_mid;
        else return i_mid;
        if ((i2 - i1) <= 1) return i1;
    }
}

/* Source.sampleKinetic (Source.java:167-198) over UniformSource.sampleParticle (sources/UniformSource.java:56-72) with
 * Spline.randomT for XY (Spline.java:582-641), Spline.pos / normal (:700-707, :947-954), LinearSegment.pos / normal
 * (LinearSegment.java:94-101, :20-45), the 1e-6*dt nudge off the surface (Source.java:186-188) and
 * DomainModule.getMesh (DomainModule.java:106-117: first mesh that strictly contains the point, else the first that
 * contains it within FLT_EPS).  Outputs the sampled particles in order and the mesh each one lands in (-1: dropped). */
/* LinearSegment.area(t), LinearSegment.java:50-80: swept area up to t (XY: t*length; RZ / ZR: lateral area of the conical frustum) */
static double seg_area(const sfo_spline *s, int i, double t, int domain_type)
{
    if (domain_type == SFO_XY) return t * s->area[i]; /* length = area(1) */
    const double px = s->x1[i] + t * (s->x2[i] - s->x1[i]), py = s->y1[i] + t * (s->y2[i] - s->y1[i]);
    double r1, z1, r2, z2;
    if (domain_type == SFO_RZ) { r1 = s->x1[i]; z1 = s->y1[i]; r2 = px; z2 = py; }
    else { r1 = s->y1[i]; z1 = s->x1[i]; r2 = py; z2 = px; }
    const double dr = r1 - r2, dz = z1 - z2;
    double A = M_PI * (r1 + r2) * sqrt(dr * dr + dz * dz);
    if (A < 0) A *= -1.0;
    return A;
}

/* Spline.randomT, Spline.java:582-641: XY takes the area fraction; axisymmetric domains search the t that sweeps the wanted
 * area with <= 10 secant steps (quirk kept: when the first guess is already within tolerance the result is x0 + (f_goal - f0)) */
static double spline_random_t(const sfo_spline *s, uint64_t *rng_state, int domain_type)
{
    const double A1 = sfo_java_next_double(rng_state) * s->spline_area;
    const int i = vec_binary_search(s->cum_area, s->n_seg + 1, A1);
    const double area = s->area[i];
    double frac = (A1 - s->cum_area[i]) / area;
    if (domain_type != SFO_XY) {
        enum { max_steps = 10 };
        const double tol = 1e-6;
        double x[max_steps], f[max_steps];
        const double f_goal = frac * area;
        x[0] = frac;
        f[0] = seg_area(s, i, x[0], domain_type);
        double diff = fabs(f[0] - f_goal) / area;
        int k = 1;
        x[1] = x[0] + (f_goal - f[0]);
        f[1] = 0;
        if (diff > tol) f[1] = seg_area(s, i, x[1], domain_type);
        while (diff > tol && k < max_steps - 1) {
            x[k + 1] = (x[k] - x[k - 1]) * (f_goal - f[k - 1]) / (f[k] - f[k - 1]) + x[k - 1];
            f[k + 1] = seg_area(s, i, x[k + 1], domain_type);
            diff = fabs(f[k + 1] - f_goal) / area;
            k++;
        }
        frac = x[k];
    }
    return i + frac;
}

void sfo_uniform_source(const sfo_spline *s, int cold_beam, double v_drift, double dt, int64_t num_mp, uint64_t *rng_state,
                        const sfo_mesh *meshes, int n_meshes, double *x, double *y, double *z, double *u, double *v,
                        double *w, int32_t *mesh_of)
{
    for (int64_t q = 0; q < num_mp; q++) {
        const double t = spline_random_t(s, rng_state, n_meshes > 0 ? meshes[0].domain_type : SFO_XY);
        int si = j2i(t); /* Spline.pos */
        double seg_t = t - si;
        if (si > s->n_seg - 1) { si = s->n_seg - 1; seg_t = 1.0; }
        double pos[3] = {s->x1[si] + seg_t * (s->x2[si] - s->x1[si]), s->y1[si] + seg_t * (s->y2[si] - s->y1[si]), 0.0};
        int sn = j2i(t); /* Spline.normal */
        if (sn > s->n_seg - 1) sn = s->n_seg - 1;
        const double n[3] = {s->nx[sn], s->ny[sn], 0.0};
        double vel[3];
        for (int k = 0; k < 3; k++) vel[k] = n[k] * v_drift; /* UniformSource.java:68 */
        if (cold_beam) vel[2] = 0; /* ColdBeamSource.java:69-71: the same sampling, vel[2] = 0 written out */
        for (int k = 0; k < 3; k++) pos[k] += vel[k] * 1e-6 * dt;
        int found = -1;
        for (int m = 0; m < n_meshes && found < 0; m++) { /* containsPosStrict, UM:164-171 */
            const sfo_mesh *mm = &meshes[m];
            const double xd0 = mm->x0[0] + (mm->ni - 1) * mm->dh[0], xd1 = mm->x0[1] + (mm->nj - 1) * mm->dh[1];
            if (pos[0] >= mm->x0[0] && pos[0] < xd0 && pos[1] >= mm->x0[1] && pos[1] < xd1) found = m;
        }
        for (int m = 0; m < n_meshes && found < 0; m++) { /* containsPos, MESH:1476-1483 */
            const sfo_mesh *mm = &meshes[m];
            double li, lj;
            sfo_xtol(mm, pos[0], pos[1], &li, &lj);
            if (!(li < -FLT_EPS || lj < -FLT_EPS || li > (mm->ni - 1 + FLT_EPS) || lj > (mm->nj - 1 + FLT_EPS))) found = m;
        }
        x[q] = pos[0]; y[q] = pos[1]; z[q] = pos[2];
        u[q] = vel[0]; v[q] = vel[1]; w[q] = vel[2];
        mesh_of[q] = found;
    }
}
