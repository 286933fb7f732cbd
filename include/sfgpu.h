/*
 * sfgpu.h -- C ABI of libstarfish_gpu.so: the B200 (sm_100a) replacement for the kinetic
 * particle hot path of particleincell/Starfish.
 *
 * The reference has no native boundary: the path lives behind the Java plugin API
 * (MaterialsModule.registerMaterialType, MaterialsModule.java:91-95; a KineticMaterial subclass
 * overriding updateFields(), KineticMaterial.java:117 -- the GrainMaterial precedent,
 * plugins/surface_processing/GrainMaterial.java:14-24).  Each entry point below names the Java
 * member(s) it replaces ("KM" = src/starfish/core/materials/KineticMaterial.java,
 * "F2D" = core/domain/Field2D.java, "MESH" = core/domain/Mesh.java, "UM" = UniformMesh.java).
 * INTEGRATION.md shows the JNI / FFM binding a Starfish maintainer adds on the Java side.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SFGPU_E* code otherwise; the message is
 *    available from sfgpu_last_error(); nothing throws across the ABI;
 *  - all pointers are caller-owned HOST memory valid for the duration of the call, the library
 *    owns all device memory; "nullable" arguments may be NULL;
 *  - fields are Java double[ni][nj] flattened row by row: index i*nj + j;
 *  - one caller thread per context (Starfish's main loop is single threaded, Starfish.java:77-121);
 *  - there is no CPU fallback: without a CUDA device sfgpu_create fails.
 */
#ifndef SFGPU_H
#define SFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFGPU_ABI_VERSION 1

/* error codes */
#define SFGPU_OK 0
#define SFGPU_EINVAL (-1)   /* bad argument                               */
#define SFGPU_ECUDA (-2)    /* CUDA runtime / launch failure               */
#define SFGPU_ENOMEM (-3)   /* device or host allocation failed            */
#define SFGPU_ENCCL (-4)    /* NCCL missing or failed                      */
#define SFGPU_EOVERFLOW (-5)/* an internal list overflowed (hand-off/slow) */
#define SFGPU_ESTATE (-6)   /* call not valid in the current state         */

/* DomainModule.DomainType, DomainModule.java:28 */
#define SFGPU_XY 0
#define SFGPU_RZ 1
#define SFGPU_ZR 2

/* Mesh.Face.val(), MESH:107-118: order of every [4] face array below */
#define SFGPU_FACE_RIGHT 0
#define SFGPU_FACE_TOP 1
#define SFGPU_FACE_LEFT 2
#define SFGPU_FACE_BOTTOM 3

/* Mesh.DomainBoundaryType.value(), MESH:140-155 */
#define SFGPU_BC_OPEN (-1)
#define SFGPU_BC_DIRICHLET 0
#define SFGPU_BC_NEUMANN 1
#define SFGPU_BC_PERIODIC 2
#define SFGPU_BC_SYMMETRY 3
#define SFGPU_BC_MESH 4
#define SFGPU_BC_SINK 5
#define SFGPU_BC_CIRCUIT 6

/* per-segment outcome of a surface hit for sfgpu_mesh_set_segments */
#define SFGPU_SURFACE_REMOVE 0
#define SFGPU_SURFACE_NONE 1
#define SFGPU_SURFACE_SPECULAR 2

/* sfgpu_inject flags */
#define SFGPU_INJECT_REWIND 1u      /* apply the -0.5*dt velocity rewind of addParticle, KM:776-794 */
#define SFGPU_INJECT_DEPOSIT_NOW 2u /* particle was already moved this step (slow-path survivor):
                                       add it to this step's deposit, move it from the next step on */
#define SFGPU_INJECT_TRANSFER 4u    /* particle enters a mesh's transfer list (KM:1377-1380): it is
                                       moved by the transfer sweeps of the next sfgpu_step.
                                       Known difference: a slow-path survivor that crosses a MESH face while the HOST
                                       finishes its sub-steps comes back this way and so finishes its residual dt (and
                                       deposits) one step later than the reference, which sweeps it in the same
                                       updateFields() (KM:131-142).  Only multi-mesh cases whose surface models stay on the
                                       host are affected; segments registered with sfgpu_mesh_set_segments are not. */

/* sfgpu_step flags */
#define SFGPU_STEP_GENERIC 1u  /* cross-check mode: no cell sort, no warp tiles; every particle gathers
                                  and deposits straight from / to global memory (FP64 REDs)         */
#define SFGPU_STEP_DEFER_FINISH 2u /* do not close the step: the host still has slow-path survivors to
                                      re-inject (SFGPU_INJECT_DEPOSIT_NOW); it calls sfgpu_finish_step.  One deferred
                                      step per context at a time: stepping or injecting into ANOTHER species before
                                      sfgpu_finish_step is refused (the step counters are per context) */
#define SFGPU_STEP_INPLACE 4u  /* force the in-place tiled step kernel + periodic cell sort (sf_fast.cuh)            */
#define SFGPU_STEP_STREAM 8u   /* force the streaming step kernel that re-sorts the store as it writes (sf_stream.cuh).
                                  Without either flag the context default applies (env SFGPU_PATH=tiled|stream).    */

/* index of each raw per-step deposit field inside the packed device buffer / sfgpu_get_deposit */
#define SFGPU_F_DEN 0 /* += mpw            KM:184, KM:1590 */
#define SFGPU_F_U 1   /* += mpw*vel[0]     KM:185, KM:1584 */
#define SFGPU_F_V 2   /* += mpw*vel[1]     KM:186, KM:1585 */
#define SFGPU_F_W 3   /* += mpw*vel[2]     KM:187, KM:1586 */
#define SFGPU_F_UU 4  /* += mpw*vel[0]^2   KM:1587         */
#define SFGPU_F_VV 5  /* += mpw*vel[1]^2   KM:1588         */
#define SFGPU_F_WW 6  /* += mpw*vel[2]^2   KM:1589         */
#define SFGPU_F_MPC 7 /* += 1 per particle in cell ((int)lc0,(int)lc1), KM:1593; a particle whose (int)lc falls outside
                         [0,ni) x [0,nj) is not counted (Java would throw ArrayIndexOutOfBounds there) */
#define SFGPU_NFIELDS 8

typedef struct sfgpu_ctx sfgpu_ctx;

/* SoA view of particles on the host, used by inject / download / take_slowpath.
 * Mirrors KineticMaterial.Particle, KM:1207-1217.  All arrays have n entries. */
typedef struct sfgpu_particles {
    int64_t n;
    double *x, *y, *z;     /* pos[0..2]                                   */
    double *u, *v, *w;     /* vel[0..2]                                   */
    double *mpw;           /* macroparticle weight                        */
    double *li, *lj;       /* lc[0..1]; nullable on inject => XtoL(pos) with the plus-edge clamp KM:762-773 */
    double *dt;            /* remaining dt; nullable on inject => 0       */
    int32_t *id;           /* nullable on inject => part_id_counter++ (KM:797) */
    int32_t *born_it;      /* nullable on inject => 0                     */
} sfgpu_particles;

/* pre-substep state handed back with slow-path particles: the arguments of
 * ProcessBoundary(part, mesh, old, lc_old), KM:471 */
typedef struct sfgpu_slow_extra {
    double *old_x, *old_y;   /* old[0..1]              */
    double *old_li, *old_lj; /* lc_old[0..1]           */
    int32_t *bounces;        /* substeps already taken */
    int32_t *mesh;           /* mesh the particle is in */
} sfgpu_slow_extra;

/* ---- lifecycle ------------------------------------------------------------------------- */
int sfgpu_abi_version(void);
/* device: CUDA ordinal; domain_type: SFGPU_XY/RZ/ZR (Starfish.getDomainType()) */
int sfgpu_create(int device, int domain_type, sfgpu_ctx **out);
void sfgpu_destroy(sfgpu_ctx *ctx);
/* ctx may be NULL: returns the last error of the calling thread's most recent failing call */
const char *sfgpu_last_error(sfgpu_ctx *ctx);

/* page-locked host memory: field / moment buffers allocated here are copied by DMA without staging (Java:
 * wrap the pointer in a direct ByteBuffer / MemorySegment).  Any other host pointer works too, through a staged copy. */
int sfgpu_host_alloc(size_t bytes, void **out);
void sfgpu_host_free(void *p);

/* ---- mesh and fields (replaces the reads of UniformMesh / Mesh state on the path) ------- */
/* UM:33-43 geometry; bc[f]: DomainBoundaryType per node of face f (MESH:215), nj entries for
 * RIGHT/LEFT, ni for TOP/BOTTOM; nbr[f] (nullable): 2 neighbour mesh ids per node or -1
 * (MeshBoundaryData.neighbor, MESH:160-165); has_seg (nullable): 1 where node[i][j].segments holds
 * a DIRICHLET or SINK segment (KM:508-518); node_vol (nullable): mesh.node_vol.data (F2D:418-431). */
int sfgpu_mesh_add(sfgpu_ctx *ctx, int32_t ni, int32_t nj, const double x0[2], const double dh[2],
                   const int8_t *const bc[4], const int32_t *const nbr[4], const uint8_t *has_seg,
                   const double *node_vol, int32_t *mesh_id);
/* surface hits on the device.  The DIRICHLET / SINK LinearSegments (boundaries/LinearSegment.java) of node[i][j].segments
 * (Mesh.setNodeControlVolumes, MESH:1215-1290) as a CSR over nodes i*nj+j (node_offs: ni*nj+1 entries, node_ids: segment indices), and per
 * segment what Material.performSurfaceInteraction (Material.java:279-300) does to THIS material's particles: SFGPU_SURFACE_REMOVE (0) = the
 * particle is removed (no interaction listed for the pair, or ABSORB, SurfaceInteraction.java:92-101), SFGPU_SURFACE_NONE (1) = it lives on
 * unchanged (NONE, :82-90, or a boundary without a material, KM:586), SFGPU_SURFACE_SPECULAR (2) = SurfaceImpactSpecular without a species
 * change (:104-149): vel[0..1] += LinearSegment.normal * (Vec.mag2(vel) * sqrt 2), the particle lives on and the hit carries the new velocity;
 * sink[k] != 0: the boundary is a SINK (KM:593-594).  With the table set, the segment part of ProcessBoundary (KM:482-603: nearest
 * LinearSegment.intersect, start-of-step exclusion, 0.9999 back-off, dt_rem) runs inside sfgpu_step and such particles no longer come back
 * through sfgpu_take_slowpath; hits are listed for sfgpu_take_surface_hits.  Models that draw random numbers (DIFFUSE / COSINE, sputtering,
 * species change) keep the host path.  The choice is per MESH: the node table replaces has_seg of sfgpu_mesh_add, so a mesh registers ALL of
 * its DIRICHLET / SINK segments here (every one of them with a deterministic outcome for this material) or none of them (and then keeps
 * has_seg + sfgpu_take_slowpath).  n_seg = 0 clears the table; the nodes stay flagged, so their particles come back through
 * sfgpu_take_slowpath again. */
int sfgpu_mesh_set_segments(sfgpu_ctx *ctx, int32_t mesh_id, int32_t n_seg, const double *x1, const double *y1, const double *x2,
                            const double *y2, const int32_t *kind, const int32_t *sink, const int32_t *node_offs, const int32_t *node_ids);
/* the surface hits of the last sfgpu_step in any order: mesh, segment index, t along the segment, velocity at impact, mpw, survived?
 * (the arguments of addSurfaceMomentum / addSurfaceMassDeposit and the boundary_charge sum, KM:586-602).  Every array nullable;
 * *n = hits of the step (min(*n, max) copied), *n_absorbed = particles the surfaces removed. */
int sfgpu_take_surface_hits(sfgpu_ctx *ctx, int32_t sp, int64_t max, int32_t *mesh, int32_t *seg, double *t, double *u, double *v,
                            double *w, double *mpw, int8_t *alive, int64_t *n, int64_t *n_absorbed);

/* efi/efj/bfi/bfj of MeshData, KM:1319-1322; bfi/bfj nullable => zero field.  Page-locked sources (sfgpu_host_alloc) are copied
 * asynchronously in stream order: leave them unchanged until the next sfgpu_step / sfgpu_sync returns. */
int sfgpu_set_fields(sfgpu_ctx *ctx, int32_t mesh_id, const double *efi, const double *efj,
                     const double *bfi, const double *bfj);

/* ---- species (one KineticMaterial) ------------------------------------------------------ */
/* q_over_m = charge/mass as Material.java:711 */
int sfgpu_species_add(sfgpu_ctx *ctx, double charge, double mass, int64_t capacity_hint, int32_t *sp);

/* KineticMaterial.addParticle(MeshData, Particle), KM:759-802 (+ MeshData.addParticle finite-velocity
 * filter KM:1356-1361).  dt_step is Starfish.getDt() for the rewind.  n_added (nullable) returns how
 * many were accepted. */
int sfgpu_inject(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, const sfgpu_particles *p, double dt_step,
                 uint32_t flags, int64_t *n_added);

/* KineticMaterial.updateFields(), KM:117-163: moveParticles(false) + transfer sweeps (KM:126-142),
 * fused with the deposit of updateFields(MeshData) KM:168-188 and updateSamples KM:1570-1595, then the
 * cross-GPU sum of the deposit when a communicator is attached. */
int sfgpu_step(sfgpu_ctx *ctx, int32_t sp, double dt, uint32_t flags);
/* Source.sampleKinetic over UniformSource.sampleParticle (core/source/Source.java:167-198,
 * sources/UniformSource.java:56-72) sampled on the device, for a Boundary of linear segments (XY; RZ / ZR with the secant search of Spline.java:594-637).  The caller passes
 * the Spline as the Java object holds it (segment end points, LinearSegment.normal, Segment.area, cum_area, spline_area) and the
 * 48-bit internal state of the java.util.Random behind Starfish.rnd(); particle p uses draws 2p, 2p+1 of that stream, so the
 * particles are the ones the sequential Java loop would create, and *rng_state returns advanced by 2*num_mp draws.  Each particle
 * gets pos = spline.pos(t) + vel*1e-6*dt, vel = normal*v_drift, mpw (the material's spwt0), born_it, lands in the mesh
 * DomainModule.getMesh picks (dropped when none contains it) and then goes through addParticle(md, part) (KM:759-802: XtoL,
 * -0.5dt rewind, id = part_id_counter++). */
typedef struct sfgpu_spline {
    int32_t n_seg;
    const double *x1, *y1, *x2, *y2; /* LinearSegment end points, n_seg each           */
    const double *nx, *ny;           /* LinearSegment.normal[0..1]                     */
    const double *area;              /* Segment.area = area(1): XY length, RZ / ZR frustum area */
    const double *cum_area;          /* n_seg + 1 entries, cum_area[0] = 0             */
    double spline_area;
} sfgpu_spline;
#define SFGPU_SOURCE_COLD_BEAM 1u /* ColdBeamSource.sampleParticle (sources/ColdBeamSource.java:56-76): the same sampling, vel[2] = 0 */
int sfgpu_source_uniform(sfgpu_ctx *ctx, int32_t sp, const sfgpu_spline *spline, uint32_t flags, double v_drift, double mpw, int32_t born_it,
                         int64_t num_mp, double dt_step, uint64_t *rng_state, int64_t *n_added);

/* closes a step opened with SFGPU_STEP_DEFER_FINISH: cross-GPU sum of the deposit and of the mover sums */
int sfgpu_finish_step(sfgpu_ctx *ctx, int32_t sp);

/* raw per-step sums, each ni*nj, any pointer nullable: order SFGPU_F_* */
int sfgpu_get_deposit(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS]);
/* running velocity-moment sums of updateSamples (KM:1570-1595), accumulated on the device at the end of every step:
 * order SFGPU_F_* = count-sum, u-sum, v-sum, w-sum, uu-sum, vv-sum, ww-sum, mpc-sum; any pointer nullable; out
 * itself nullable to read only num_samples (KM:1557).  sfgpu_clear_samples = clearSamples(), KM:1509-1528. */
int sfgpu_get_samples(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS], int64_t *num_samples);
int sfgpu_clear_samples(sfgpu_ctx *ctx, int32_t sp);
/* nd,u,v,w after KM:190-196 (U,V,W /= Den; Den /= node_vol); needs node_vol; any nullable */
int sfgpu_get_moments(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, double *nd, double *u, double *v,
                      double *w);
/* sums5 = {N, Px, Py, Pz, E} of KM:406-413 (not yet multiplied by mass, KM:252-258), summed over ranks
 * when a communicator is attached; counts are local to this context */
int sfgpu_get_sums(sfgpu_ctx *ctx, int32_t sp, double sums5[5], int64_t *np_alive, int64_t *n_exited,
                   int64_t *n_slow);
/* MeshData.getNp(), KM:1385-1390 (mesh_id <0: KineticMaterial.getNp(), KM:1297) */
int sfgpu_np(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t *np);

/* particles that need the unchanged Java ProcessBoundary (KM:471-750): a DIRICHLET/SINK segment lies in
 * the node bounding box of their substep, or a CIRCUIT face was hit by an electron.  They are handed over
 * in their pre-ProcessBoundary state and removed from the device store; survivors come back through
 * sfgpu_inject(.., SFGPU_INJECT_DEPOSIT_NOW).  *n returns the number copied (<= max). */
int sfgpu_take_slowpath(sfgpu_ctx *ctx, int32_t sp, int64_t max, sfgpu_particles *out, sfgpu_slow_extra *extra,
                        int64_t *n);

/* iterators / output / restart / collisions (KM:271, :1187, :904-1000): copy particles [first, first+n)
 * of a mesh's store to the host, or overwrite them after a host-side mutation */
int sfgpu_download(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t first, sfgpu_particles *out);
int sfgpu_upload(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t first, const sfgpu_particles *in);

/* restart.bin, particle section of one mesh (KineticMaterial.saveRestartData / loadRestartData, KM:904-1000): the
 * exact DataOutputStream bytes -- long np, then per particle pos/vel interleaved, lc, dt, mpw, mass, born_it, id, all
 * big endian (96 bytes per particle), packed on the device.  save: buf NULL => only *bytes_needed is returned.
 * load: goes through addParticle(md, part) like the reference (caller-supplied lc, -0.5dt rewind re-applied, ids
 * re-assigned); *bytes_used tells the caller where the field section of the stream starts. */
int sfgpu_restart_save(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, void *buf, int64_t buf_bytes, int64_t *bytes_needed);
int sfgpu_restart_load(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, const void *buf, int64_t buf_bytes, double dt_step,
                       int64_t *bytes_used, int64_t *n_loaded);

/* explicit cell sort + compaction (sortParticlesToCells, KM:1150-1179: order only, no result change) */
int sfgpu_sort(sfgpu_ctx *ctx, int32_t sp);
/* SURVEY 8f-2, per-cell particle lists for consumers that bin by cell (collisions/DSMC.java:194-252, MCC.java:167-216; the reference's own
 * sortParticlesToCells, KM:1150-1179): cell-sorts the store of one mesh and returns, for every CELL c = i*(nj-1) + j, the index of its first particle in
 * sfgpu_download / sfgpu_upload order (cell_first, (ni-1)*(nj-1) entries) and its population (cell_count).  Indices [0, *n_sorted) are covered; the few
 * exceptional records (stale lc after a boundary clamp, residual dt) follow unsorted at [*n_sorted, np).  Valid until the next call that moves particles. */
int sfgpu_cell_lists(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, int64_t *cell_first, int32_t *cell_count, int64_t *n_sorted);
/* sfgpu_step re-sorts the store by cell every `steps` steps (default 3, env SFGPU_SORT_EVERY).
 * KineticMaterial.mergeParticles (KM:1008-1143, particle_merge_skip > 0) is NOT offered: materials that merge keep type="kinetic". */
int sfgpu_set_sort_interval(sfgpu_ctx *ctx, int32_t steps);
/* halo of the tiled kernel's warp-private accumulation tile, in cells around the 4 x 4-cell tile a work item covers.  1: the deposits of a particle must land
 * within one cell of the cell it was sorted into (true while |v| dt < ~0.3 cells; the sort key is the predicted cell) -- smaller tile, cheaper flush;
 * 2: two cells.  A deposit outside the tile is still made (k_fast_deferred, global atomics), only slower: results never depend on it.
 * 0 (default, env SFGPU_HALO): start with 1 and switch to 2 for good once a step leaves more than 0.5 % of its deposits to the deferred kernel. */
int sfgpu_set_tile_halo(sfgpu_ctx *ctx, int32_t halo);
int sfgpu_get_tile_halo(sfgpu_ctx *ctx, int32_t *halo, int32_t *automatic);

/* ---- multi GPU: particles are partitioned over contexts, meshes replicated ------------- */
/* 128-byte NCCL unique id, created on one rank and distributed by the host */
int sfgpu_comm_unique_id(void *id128);
int sfgpu_comm_init(sfgpu_ctx *ctx, int32_t nranks, int32_t rank, const void *id128);
/* device pointer + element count of the packed [SFGPU_NFIELDS][ni][nj] deposit buffer, for hosts that
 * run their own collective */
int sfgpu_deposit_device_ptr(sfgpu_ctx *ctx, int32_t sp, int32_t mesh_id, void **ptr, int64_t *count);

/* ---- single-caller multi GPU ------------------------------------------- */
/* Starfish's main loop is one thread (Starfish.java:77-121).  A group owns one context per GPU and a worker thread per context:
 * every call below fans out to all GPUs concurrently and returns when all are done, so the single Java thread drives 8 GPUs
 * without deadlocking on the NCCL all-reduce.  Meshes, segments and fields are replicated; injected particles are partitioned
 * by index (contiguous, balanced; ids assigned from one counter so they stay unique); sfgpu_multi_step = sfgpu_step everywhere
 * with the deposit and the mover sums all-reduced inside; results are read back once, from rank 0.  device_ids NULL: 0..n-1.
 * Per-GPU entry points (slow path, download / upload, restart, surface hits) remain available through sfgpu_multi_ctx(g, rank). */
typedef struct sfgpu_multi sfgpu_multi;
int sfgpu_multi_create(int32_t n, const int32_t *device_ids, int domain_type, sfgpu_multi **out);
void sfgpu_multi_destroy(sfgpu_multi *g);
int32_t sfgpu_multi_size(sfgpu_multi *g);
sfgpu_ctx *sfgpu_multi_ctx(sfgpu_multi *g, int32_t rank);
const char *sfgpu_multi_last_error(sfgpu_multi *g);
int sfgpu_multi_mesh_add(sfgpu_multi *g, int32_t ni, int32_t nj, const double x0[2], const double dh[2], const int8_t *const bc[4],
                         const int32_t *const nbr[4], const uint8_t *has_seg, const double *node_vol, int32_t *mesh_id);
int sfgpu_multi_mesh_set_segments(sfgpu_multi *g, int32_t mesh_id, int32_t n_seg, const double *x1, const double *y1, const double *x2,
                                  const double *y2, const int32_t *kind, const int32_t *sink, const int32_t *node_offs, const int32_t *node_ids);
int sfgpu_multi_set_fields(sfgpu_multi *g, int32_t mesh_id, const double *efi, const double *efj, const double *bfi, const double *bfj);
int sfgpu_multi_species_add(sfgpu_multi *g, double charge, double mass, int64_t capacity_hint, int32_t *sp);
int sfgpu_multi_inject(sfgpu_multi *g, int32_t sp, int32_t mesh_id, const sfgpu_particles *p, double dt_step, uint32_t flags, int64_t *n_added);
int sfgpu_multi_step(sfgpu_multi *g, int32_t sp, double dt, uint32_t flags);
int sfgpu_multi_finish_step(sfgpu_multi *g, int32_t sp);
int sfgpu_multi_get_moments(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *nd, double *u, double *v, double *w);
int sfgpu_multi_get_deposit(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS]);
int sfgpu_multi_get_samples(sfgpu_multi *g, int32_t sp, int32_t mesh_id, double *const out[SFGPU_NFIELDS], int64_t *num_samples);
int sfgpu_multi_clear_samples(sfgpu_multi *g, int32_t sp);
int sfgpu_multi_get_sums(sfgpu_multi *g, int32_t sp, double sums5[5], int64_t *np_alive, int64_t *n_exited, int64_t *n_slow);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* device milliseconds of the last step, from CUDA events on the context's own stream: ms_total spans the
 * whole step (the periodic cell sort when one ran, memsets, kernels, collective, running sums), ms_kernel only the fused
 * move+deposit kernel(s) of the main pass; launches = kernels launched by the step.  Any pointer nullable. */
int sfgpu_last_step_timing(sfgpu_ctx *ctx, float *ms_total, float *ms_kernel, int32_t *launches);
/* which step kernel ran in the last sfgpu_step: 0 = tiled in-place kernel (the default; the store is re-sorted by cell every
 * few steps, KM:1150-1179 being the closest reference member), 1 = streaming kernel (moves, deposits and re-sorts in one pass;
 * SFGPU_STEP_STREAM or SFGPU_PATH=stream), 2 = generic (SFGPU_STEP_GENERIC). */
int sfgpu_last_step_kernel(sfgpu_ctx *ctx, int32_t *kind);
/* diagnostics of the last step: particles whose deposit missed the warp tile of their sort position and went
 * through global atomics instead (grows between cell sorts; the step re-sorts early when it passes 1/64) */
int sfgpu_last_step_counters(sfgpu_ctx *ctx, int64_t *n_fallback);
/* cudaStreamSynchronize on the context's stream */
int sfgpu_sync(sfgpu_ctx *ctx);
/* bracket any sequence of calls with CUDA events on the context's stream (the stream every kernel of the
 * library is launched on): stop synchronises and returns the device milliseconds since start */
int sfgpu_timer_start(sfgpu_ctx *ctx);
int sfgpu_timer_stop(sfgpu_ctx *ctx, float *ms);
/* kernels launched by this context since sfgpu_create (cumulative) */
int sfgpu_launch_count(sfgpu_ctx *ctx, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* SFGPU_H */
