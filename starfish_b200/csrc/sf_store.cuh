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
    unsigned long long n_slow;    // slow-path list length
    unsigned long long n_xfer_copies;
    unsigned long long overflow;  // a list ran out of room
    unsigned long long n_bad;     // non-finite velocity on inject (KM:1357-1361)
    unsigned long long xfer_n[16]; // per-mesh transfer list length (SF_MAX_MESHES)
};
#define SF_MAX_MESHES 16
