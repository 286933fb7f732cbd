// sf_store.cuh -- device-side layouts shared by the kernels and the host API.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Full particle record, structure of arrays: KineticMaterial.Particle (KM:1207-1217).  The generic kernels work on
// these; the tiled store keeps only x..mpw + tag per particle and re-derives lc (see DESIGN.md "Data layout").
struct RecPtrs {
    double *x, *y, *z, *u, *v, *w, *mpw, *li, *lj, *dt;
    int2 *tag; // {id, born_it}
};
#define SF_REC_NDOUBLES 10

// Fast store: what a NORMAL particle needs between steps (lc == XtoL(pos) and dt == 0 are implied, so they are
// not stored): 7 doubles = 56 B read per push, 6 written back in place -> the 104 algorithmic bytes of SURVEY 8d.
// tag {id, born_it} rides along but is only touched by the sort.  A slot with mpw == NaN is vacant.
struct FastPtrs {
    double *x, *y, *z, *u, *v, *w, *mpw;
    int2 *tag;
};

// one unit of work of the tiled step kernel: a run of the cell-sorted fast store that lies in one tile
struct WorkItem {
    unsigned long long begin;
    int count;
    int tile; // ti * ntj + tj
};

#ifndef SF_TILE
#define SF_TILE 4       // cells per tile edge.  Measured on config B (kernel ms): 8 -> 0.68 (15 warps / SM fit), 6 -> 0.64 (18), 4 -> 0.62 (20, the
                        // register limit), 2 -> 0.91 (the per-item tile flush dominates); 5 makes the node rows 8 doubles long: bank conflicts, 0.78
#endif
#ifndef SF_HALO
#define SF_HALO 2       // extra cells kept around the tile in the warp-private accumulation tile
#endif
#define SF_NT (SF_TILE + 2 * SF_HALO + 1) // nodes per edge of the accumulation tile
#ifndef SF_ITEM_MAX
#define SF_ITEM_MAX 2048 // particles per work item
#endif

// slow-path hand-over list: record + ProcessBoundary arguments
struct SlowPtrs {
    RecPtrs rec;
    double *old_x, *old_y, *old_li, *old_lj;
    int *bounces, *mesh;
    unsigned long long cap;
};

// per-mesh transfer list (MeshData.transfer_particles, KM:1346) as seen by kernels
struct XferDev {
    RecPtrs rec;
    unsigned long long cap;
};

// counters of one sfgpu_step, zeroed at its start
struct StepCounters {
    double sums[5]; // N, Px, Py, Pz, E  (KM:406-413)
    unsigned long long n_out[16]; // per-mesh survivors appended to the next-step list
    unsigned long long n_exited;  // SF_DEAD
    unsigned long long n_removed; // SF_REMOVED
    unsigned long long n_absorbed; // SF_ABSORBED
    unsigned long long n_hits;    // surface-hit list cursor
    unsigned long long n_slow;    // slow-path list length
    unsigned long long n_xfer_copies;
    unsigned long long overflow;  // a list ran out of room
    unsigned long long n_bad;     // non-finite velocity on inject (KM:1357-1361)
    unsigned long long xfer_n[16]; // per-mesh transfer list length (SF_MAX_MESHES)
    unsigned long long n_exc[16];  // per-mesh: records appended by k_inject_fast / k_records_to_fast
    unsigned long long fast_n[16]; // per-mesh: fast-store length cursor (records that became normal are appended)
    unsigned long long n_fallback; // deposits that missed the warp tile and went to global atomics
    unsigned long long n_flush;    // segment flushes of the tiled kernel (diagnostic)
    unsigned long long n_defer[16]; // per mesh: particles k_fast_step left to k_fast_deferred
    long long fast_delta[16];      // per-mesh change of the number of live fast-store particles
    unsigned int queue[16];        // per-mesh work-item queue head of the tiled kernel
};
#define SF_MAX_MESHES 16
