/*
 * sf_oracle.h -- CPU ORACLE for the Starfish kinetic-particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's Java
 * algorithm (KineticMaterial move + deposit + sampling).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (libstarfish_gpu.so / starfish_b200) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 *  and no JVM exists in the build image, so the oracle cannot be
 * checked against reference outputs.  It is pinned instead by analytic known-answer tests
 * (tests/test_oracle_kat.py) derived from the Java source.
 *
 * Java semantics reproduced: strict left-to-right FP64, no FMA contraction (compile with
 * -ffp-contract=off), (int) casts truncate toward zero / saturate / NaN->0, fields are
 * double[ni][nj] flattened as i*nj+j.
 *
 * Reference files restated (paths under src/starfish/core/):
 *   materials/KineticMaterial.java  (KM)   domain/Field2D.java (F2D)
 *   domain/UniformMesh.java (UM)           domain/Mesh.java (MESH)     common/Vec.java
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DomainType, DomainModule.java:28 */
enum { SFO_XY = 0, SFO_RZ = 1, SFO_ZR = 2 };
/* Mesh.Face values, MESH:107-118 */
enum { SFO_RIGHT = 0, SFO_TOP = 1, SFO_LEFT = 2, SFO_BOTTOM = 3 };
/* Mesh.DomainBoundaryType values, MESH:140-155 */
enum { SFO_OPEN = -1, SFO_DIRICHLET = 0, SFO_NEUMANN = 1, SFO_PERIODIC = 2, SFO_SYMMETRY = 3,
       SFO_MESH = 4, SFO_SINK = 5, SFO_CIRCUIT = 6 };

/* per-particle outcome of one mover pass */
enum {
    SFO_ALIVE = 0,     /* still in this mesh                                               */
    SFO_REMOVED = 1,   /* mpw<=0 on entry, KM:322                                          */
    SFO_DEAD = 2,      /* left through OPEN / default face, or CIRCUIT ion                 */
    SFO_SLOW = 3,      /* needs the Java slow path: a DIRICHLET/SINK segment is attached to
                          a node of the move's bounding box (KM:504-603) or CIRCUIT electron */
    SFO_TRANSFER = 4,  /* crossed a MESH face, KM:708-722; dead here, copies listed        */
    SFO_ABSORBED = 5   /* hit a segment whose surface model removes it, KM:586-603          */
};

/* a DIRICHLET / SINK LinearSegment of a solid Boundary as ProcessBoundary sees it (LinearSegment.java:21-47, :113-179)
 * with the outcome of Material.performSurfaceInteraction (Material.java:279-300) folded into `kind` */
typedef struct {
    double x1, y1, x2, y2;
    int32_t kind;  /* 0: the particle dies (no interaction listed, or ABSORB: SurfaceInteraction.java:92-101);
                      1: it lives on with unchanged velocity (NONE, SurfaceInteraction.java:82-90);
                      2: SPECULAR without a species change (SurfaceInteraction.java:104-149): vel[0..1] += normal * (Vec.mag2(vel) * SQRT2), alive */
    int32_t sink;  /* the Boundary is of type SINK: dies regardless, KM:593-594 */
} sfo_segment;

/* surface hits of a mover pass, in the order they happen (single thread) / any order (sfo_move_mt): what the Java
 * side needs for addSurfaceMomentum / addSurfaceMassDeposit / boundary_charge, KM:590-602 */
typedef struct {
    int64_t cap, n;
    int32_t *seg;      /* segment index                     */
    double *t;         /* position along the segment [0,1]  */
    double *u, *v, *w; /* velocity at impact                */
    double *mpw;
    int8_t *alive;
} sfo_hits;

typedef struct {
    int32_t ni, nj;            /* node counts                                  */
    double x0[2], dh[2];       /* UM:33-43                                     */
    int32_t domain_type;       /* SFO_XY / RZ / ZR                             */
    const int8_t *bc[4];       /* per-face per-node DomainBoundaryType; RIGHT/LEFT: nj entries,
                                  TOP/BOTTOM: ni entries (MESH:167, :215)        */
    const int32_t *nbr[4];     /* per-face per-node 2 neighbour mesh ids (or -1), nullable */
    const uint8_t *has_seg;    /* ni*nj, 1 where node.segments holds a DIRICHLET/SINK segment; nullable */
    /* node.segments (MESH:1215-1290) restricted to DIRICHLET / SINK segments, as CSR over nodes i*nj+j; all nullable:
     * without them a particle whose sub-step touches a has_seg node is handed back as SFO_SLOW */
    const int32_t *seg_offs;   /* ni*nj + 1 */
    const int32_t *seg_ids;
    const sfo_segment *segs;
    sfo_hits *hits;            /* nullable */
    const double *efi, *efj;   /* ni*nj each                                   */
    const double *bfi, *bfj;   /* nullable => zero field                       */
} sfo_mesh;

typedef struct {
    int64_t n;
    double *x, *y, *z;         /* Particle.pos[0..2]  KM:1208 */
    double *u, *v, *w;         /* Particle.vel[0..2]          */
    double *mpw;               /* macroparticle weight        */
    double *li, *lj;           /* Particle.lc[0..1]           */
    double *dt;                /* remaining dt                */
} sfo_particles;

/* extra per-particle outputs of a mover pass (all nullable except status) */
typedef struct {
    int8_t *status;            /* SFO_ALIVE ...                                           */
    int32_t *xfer_mask;        /* bit m set: copy goes to neighbour m of the exit node    */
    int32_t *xfer_mesh;        /* [2n] neighbour mesh ids for the set bits                */
    double *xfer_li, *xfer_lj; /* [2n] lc in the neighbour frame                          */
    double *old_x, *old_y, *old_li, *old_lj; /* pre-substep state for SFO_SLOW            */
    int32_t *bounces;          /* substeps consumed                                       */
} sfo_move_out;

/* a Boundary made of linear segments, as the Java Spline holds it (Spline.java: segments, cum_area, spline_area;
 * LinearSegment.java:20-45: normal = (-dy, dx, 0) / length) */
typedef struct {
    int32_t n_seg;
    const double *x1, *y1, *x2, *y2; /* LinearSegment end points                  */
    const double *nx, *ny;           /* LinearSegment.normal[0..1]                */
    const double *area;              /* Segment.area (XY: the segment length)     */
    const double *cum_area;          /* n_seg + 1 entries, cum_area[0] = 0        */
    double spline_area;
} sfo_spline;

uint64_t sfo_java_seed(int64_t seed);              /* java.util.Random.setSeed scrambling      */
int32_t sfo_java_next_int(uint64_t *state);        /* java.util.Random.nextInt()               */
double sfo_java_next_double(uint64_t *state);      /* java.util.Random.nextDouble()            */
void sfo_uniform_source(const sfo_spline *s, int cold_beam, double v_drift, double dt, int64_t num_mp, uint64_t *rng_state,
                        const sfo_mesh *meshes, int n_meshes, double *x, double *y, double *z, double *u, double *v,
                        double *w, int32_t *mesh_of);

/* LinearSegment.intersect(p3, p4), LinearSegment.java:113-160: t[0] along the segment, t[1] along p3->p4, (-1,-1) if none */
void sfo_segment_intersect(const sfo_segment *s, const double p3[2], const double p4[2], double t[2]);
double sfo_gather(const double *d, int ni, int nj, double fi, double fj);      /* F2D:300-350 */
double sfo_gather_safe(const double *d, int ni, int nj, double fi, double fj); /* F2D:371-390 */
void sfo_scatter(double *d, const sfo_mesh *m, double fi, double fj, double val); /* F2D:244-295 */
void sfo_xtol(const sfo_mesh *m, double x, double y, double *li, double *lj);  /* UM:154-161 */
int sfo_contains_pos(const sfo_mesh *m, double x, double y);                   /* MESH:1476-1483 */
void sfo_boris(double qm, double dtp, const double E[3], const double B[3], double vel[3]); /* KM:847-893 */
void sfo_mirror(double vel[3], const double n[3]);                             /* Vec.java:406-418 */

/* ParticleMover.run, KM:298-422, over particles [first, first+count) in order.
 * meshes[] is needed only for MESH hand-off (neighbour XtoL / containsPos).
 * sums5 += {N, Px, Py, Pz, E} of the survivors in this range (KM:406-413), unscaled by mass. */
void sfo_move(const sfo_mesh *meshes, int mesh_id, double q_over_m, double charge, double dt,
              int particle_transfer, sfo_particles *p, int64_t first, int64_t count,
              sfo_move_out *out, double sums5[5]);

/* same, the reference's threading: T ParticleMover threads over T contiguous blocks
 * (KM:200-261).  sums are reduced in block order like KM:252-258. */
void sfo_move_mt(const sfo_mesh *meshes, int mesh_id, double q_over_m, double charge, double dt,
                 int particle_transfer, sfo_particles *p, sfo_move_out *out, double sums5[5],
                 int threads);

/* KM:168-197 without the final normalisation: den,u,v,w must be zeroed by the caller */
void sfo_deposit(const sfo_mesh *m, const sfo_particles *p, double *den, double *u, double *v,
                 double *w);
/* F2D:403-414 and F2D:418-431 */
void sfo_divide_by_field(double *d, const double *by, int64_t n);
void sfo_scale_by_vol(double *d, const double *node_vol, int64_t n);
/* KM:1570-1595: running sums, accumulated in place */
void sfo_sample(const sfo_mesh *m, const sfo_particles *p, double *count_sum, double *u_sum,
                double *v_sum, double *w_sum, double *uu_sum, double *vv_sum, double *ww_sum,
                double *mpc_sum);
/* KM:759-802 for particles [first, first+count): lc (if compute_lc) + clamp, -0.5dt rewind, dt=0 */
void sfo_add_particles(const sfo_mesh *m, double q_over_m, double dt, int compute_lc,
                       sfo_particles *p, int64_t first, int64_t count);

#ifdef __cplusplus
}
#endif
#endif
