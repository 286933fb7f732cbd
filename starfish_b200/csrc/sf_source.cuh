// sf_source.cuh -- SURVEY 8f-1: particle injection by the reference's UniformSource / ColdBeamSource, sampled on the device.
//
// Restates Source.sampleKinetic (core/source/Source.java:167-198) over UniformSource.sampleParticle
// (sources/UniformSource.java:56-72) for a Boundary of linear segments (XY, RZ and ZR domains): Spline.randomT (Spline.java:582-641),
// Vec.binarySearch (Vec.java:529-546), Spline.pos / normal (:700-707, :947-954), LinearSegment.pos (LinearSegment.java:94-101),
// the 1e-6*dt nudge off the surface, DomainModule.getMesh (DomainModule.java:106-117).  The random numbers are those of
// java.util.Random: particle p of a call uses draws 2p and 2p+1 of the stream (nextDouble = next(26), next(27)), reached by
// jumping the 48-bit LCG ahead, so the result is the one the sequential Java loop produces and the host gets the advanced
// state back.  What follows (XtoL, plus-edge clamp, -0.5dt rewind, ids) is k_inject_fast = addParticle(md, part), KM:759-802.
#pragma once
#include "sf_fast.cuh"

struct SplineDev {
    int n_seg;
    const double *x1, *y1, *x2, *y2, *nx, *ny, *area, *cum_area;
    double spline_area;
};

#define SF_JAVA_A 0x5DEECE66DULL
#define SF_JAVA_C 0xBULL
#define SF_JAVA_MASK ((1ULL << 48) - 1)

// state after k steps of s -> (a*s + c) mod 2^48
__host__ __device__ inline unsigned long long sf_java_jump(unsigned long long s, unsigned long long k)
{
    unsigned long long A = 1, Cc = 0, ca = SF_JAVA_A, cc = SF_JAVA_C;
    while (k) {
        if (k & 1ULL) {
            A = (A * ca) & SF_JAVA_MASK;
            Cc = (Cc * ca + cc) & SF_JAVA_MASK;
        }
        cc = ((ca + 1) * cc) & SF_JAVA_MASK;
        ca = (ca * ca) & SF_JAVA_MASK;
        k >>= 1;
    }
    return (A * s + Cc) & SF_JAVA_MASK;
}

__device__ __forceinline__ int sf_java_next(unsigned long long &s, int bits)
{
    s = (s * SF_JAVA_A + SF_JAVA_C) & SF_JAVA_MASK;
    return (int)(s >> (48 - bits));
}

__device__ __forceinline__ double sf_java_next_double(unsigned long long &s)
{
    const long long hi = (long long)sf_java_next(s, 26), lo = (long long)sf_java_next(s, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}

// Vec.binarySearch, Vec.java:529-546
__device__ __forceinline__ int sf_vec_binary_search(const double *__restrict__ vec, int n, double val)
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

// LinearSegment.area(t), LinearSegment.java:50-80: area swept up to t (XY: t*length; RZ / ZR: lateral area of the conical frustum)
__device__ __forceinline__ double sf_seg_area(const SplineDev &s, int i, double t, int domain)
{
    if (domain == SFGPU_XY) return t * s.area[i];
    const double px = s.x1[i] + t * (s.x2[i] - s.x1[i]), py = s.y1[i] + t * (s.y2[i] - s.y1[i]);
    double r1, z1, r2, z2;
    if (domain == SFGPU_RZ) { r1 = s.x1[i]; z1 = s.y1[i]; r2 = px; z2 = py; }
    else { r1 = s.y1[i]; z1 = s.x1[i]; r2 = py; z2 = px; }
    const double dr = r1 - r2, dz = z1 - z2;
    double A = 3.141592653589793 * (r1 + r2) * sqrt(dr * dr + dz * dz); // Math.PI
    if (A < 0) A *= -1.0;
    return A;
}

// Spline.randomT, Spline.java:582-641.  XY: segment + area fraction.  Axisymmetric: <= 10 secant steps towards the t that sweeps
// the sampled area (quirk kept: an already converged first guess returns x0 + (f_goal - f0)).
__device__ __forceinline__ double sf_spline_random_t(const SplineDev &s, unsigned long long &st, int domain)
{
    const double A1 = sf_java_next_double(st) * s.spline_area;
    const int i = sf_vec_binary_search(s.cum_area, s.n_seg + 1, A1);
    const double area = s.area[i];
    double frac = (A1 - s.cum_area[i]) / area;
    if (domain != SFGPU_XY) {
        const double tol = 1e-6, f_goal = frac * area;
        double x0 = frac, f0 = sf_seg_area(s, i, x0, domain); // x[k-1], f[k-1]
        double diff = fabs(f0 - f_goal) / area;
        double x1 = x0 + (f_goal - f0), f1 = 0;               // x[k], f[k]
        if (diff > tol) f1 = sf_seg_area(s, i, x1, domain);
        for (int k = 1; diff > tol && k < 9; k++) {
            const double x2 = (x1 - x0) * (f_goal - f0) / (f1 - f0) + x0;
            const double f2 = sf_seg_area(s, i, x2, domain);
            diff = fabs(f2 - f_goal) / area;
            x0 = x1; f0 = f1; x1 = x2; f1 = f2;
        }
        frac = x1;
    }
    return i + frac;
}

// one thread per sampled particle: position, velocity, and the mesh DomainModule.getMesh() picks (-1: outside every mesh)
__global__ void __launch_bounds__(256)
k_source_uniform(SplineDev s, int domain, int cold_beam, double v_drift, double dt, unsigned long long n, unsigned long long rng_state, const MeshDev *__restrict__ meshes,
                 int n_meshes, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z, double *__restrict__ u,
                 double *__restrict__ v, double *__restrict__ w, int *__restrict__ mesh_of)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    unsigned long long st = sf_java_jump(rng_state, 2ULL * q);
    const double t = sf_spline_random_t(s, st, domain);
    int si = sf_j2i(t); // Spline.pos
    double seg_t = t - si;
    if (si > s.n_seg - 1) { si = s.n_seg - 1; seg_t = 1.0; }
    double pos[3] = {s.x1[si] + seg_t * (s.x2[si] - s.x1[si]), s.y1[si] + seg_t * (s.y2[si] - s.y1[si]), 0.0};
    int sn = sf_j2i(t); // Spline.normal
    if (sn > s.n_seg - 1) sn = s.n_seg - 1;
    const double nrm[3] = {s.nx[sn], s.ny[sn], 0.0};
    double vel[3];
#pragma unroll
    for (int k = 0; k < 3; k++) vel[k] = nrm[k] * v_drift; // UniformSource.java:68
    if (cold_beam) vel[2] = 0.0; // ColdBeamSource.java:69-71: the same sampling with vel[2] = 0 written out (+0 even for a negative drift)
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] += vel[k] * 1e-6 * dt; // Source.java:186-188
    int found = -1;
    for (int m = 0; m < n_meshes && found < 0; m++) { // containsPosStrict, UM:164-171 (xd = x0 + (n-1)*dh, UM:131-135)
        const MeshDev &mm = meshes[m];
        const double xd0 = mm.x0 + (mm.ni - 1) * mm.dhx, xd1 = mm.y0 + (mm.nj - 1) * mm.dhy;
        if (pos[0] >= mm.x0 && pos[0] < xd0 && pos[1] >= mm.y0 && pos[1] < xd1) found = m;
    }
    for (int m = 0; m < n_meshes && found < 0; m++) { // containsPos, MESH:1476-1483
        double li, lj;
        if (sf_contains_pos(meshes[m], pos[0], pos[1], li, lj)) found = m;
    }
    x[q] = pos[0]; y[q] = pos[1]; z[q] = pos[2];
    u[q] = vel[0]; v[q] = vel[1]; w[q] = vel[2];
    mesh_of[q] = found;
}

// flag = 1 where the particle goes to `mesh` (mesh >= 0) or to any mesh (mesh < 0)
__global__ void k_source_flags(const int *__restrict__ mesh_of, unsigned long long n, int mesh, unsigned *__restrict__ flag)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) flag[q] = (mesh < 0 ? mesh_of[q] >= 0 : mesh_of[q] == mesh) ? 1u : 0u;
}

// append the particles of one mesh, in sampling order, behind the fast store; ids count up over the accepted particles
__global__ void k_source_append(const int *__restrict__ mesh_of, unsigned long long n, int mesh, const unsigned *__restrict__ rank_mesh,
                                const unsigned *__restrict__ rank_all, const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, const double *__restrict__ u, const double *__restrict__ v,
                                const double *__restrict__ w, double mpw, int id_base, int born_it, FastPtrs fs, unsigned long long first)
{
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || mesh_of[q] != mesh) return;
    const size_t d = first + rank_mesh[q];
    fs.x[d] = x[q]; fs.y[d] = y[q]; fs.z[d] = z[q];
    fs.u[d] = u[q]; fs.v[d] = v[q]; fs.w[d] = w[q];
    fs.mpw[d] = mpw;
    fs.tag[d] = make_int2(id_base + (int)rank_all[q], born_it);
}
