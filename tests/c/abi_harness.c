/* abi_harness.c -- plain C caller of include/sfgpu.h: the compiler checks every prototype it uses against the header,
 * and on a GPU box the run walks create -> mesh -> fields -> species -> inject -> step -> sums / deposit / moments ->
 * download -> destroy with conservation checks.  Built and run by tests/test_abi.py (gcc, no CUDA headers needed).
 *
 *   gcc -std=c99 -Wall -Werror -I include tests/c/abi_harness.c -o abi_harness -L starfish_b200 -l:libstarfish_gpu.so -lm
 *   ./abi_harness            run on cuda:0, exit 0 = OK
 *   ./abi_harness --link     only prove that the library loads and sfgpu_abi_version() answers (CPU boxes)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sfgpu.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc_ = (call);                                                                             \
        if (rc_ != SFGPU_OK) {                                                                        \
            fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_, sfgpu_last_error(ctx)); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char **argv)
{
    sfgpu_ctx *ctx = NULL;
    if (sfgpu_abi_version() != SFGPU_ABI_VERSION) {
        fprintf(stderr, "ABI version %d, header says %d\n", sfgpu_abi_version(), SFGPU_ABI_VERSION);
        return 1;
    }
    if (argc > 1 && strcmp(argv[1], "--link") == 0) {
        printf("abi_harness link OK (ABI %d)\n", sfgpu_abi_version());
        return 0;
    }
    enum { NI = 41, NJ = 33, N = 20000 };
    const double x0[2] = {0.0, 0.0}, dh[2] = {1e-3, 2e-3}, dt = 1e-7;
    CHECK(sfgpu_create(0, SFGPU_XY, &ctx));

    /* mesh: all faces OPEN, no segments, unit node volumes */
    int8_t bc_i[NI], bc_j[NJ];
    memset(bc_i, SFGPU_BC_OPEN, sizeof bc_i);
    memset(bc_j, SFGPU_BC_OPEN, sizeof bc_j);
    const int8_t *bc[4];
    bc[SFGPU_FACE_RIGHT] = bc_j; bc[SFGPU_FACE_LEFT] = bc_j; bc[SFGPU_FACE_TOP] = bc_i; bc[SFGPU_FACE_BOTTOM] = bc_i;
    const int32_t *nbr[4] = {NULL, NULL, NULL, NULL};
    static double node_vol[NI * NJ], efi[NI * NJ], efj[NI * NJ];
    for (int k = 0; k < NI * NJ; k++) { node_vol[k] = 1.0; efi[k] = 25.0; efj[k] = -10.0; }
    int32_t mesh = -1, sp = -1;
    CHECK(sfgpu_mesh_add(ctx, NI, NJ, x0, dh, bc, nbr, NULL, node_vol, &mesh));
    CHECK(sfgpu_set_fields(ctx, mesh, efi, efj, NULL, NULL));
    CHECK(sfgpu_species_add(ctx, 1.602176565e-19, 16 * 1.660538921e-27, N, &sp));

    /* particles on a lattice strictly inside the mesh, slow enough to stay inside for the three steps */
    static double x[N], y[N], z[N], u[N], v[N], w[N], mpw[N];
    double wsum = 0;
    for (int k = 0; k < N; k++) {
        x[k] = (0.5 + (k % 200) * 0.19) * dh[0];
        y[k] = (0.5 + (k / 200) * 0.31) * dh[1];
        z[k] = 0;
        u[k] = 300.0 * ((k % 7) - 3);
        v[k] = 200.0 * ((k % 5) - 2);
        w[k] = 10.0 * (k % 3);
        mpw[k] = 1.0 + (k % 4);
        wsum += mpw[k];
    }
    sfgpu_particles p;
    memset(&p, 0, sizeof p);
    p.n = N; p.x = x; p.y = y; p.z = z; p.u = u; p.v = v; p.w = w; p.mpw = mpw;
    int64_t added = 0, np = 0, n_exited = 0, n_slow = 0, nlaunch = 0;
    CHECK(sfgpu_inject(ctx, sp, mesh, &p, dt, SFGPU_INJECT_REWIND, &added));
    if (added != N) { fprintf(stderr, "inject accepted %lld of %d\n", (long long)added, N); return 1; }

    static double den[NI * NJ], mpc[NI * NJ], nd[NI * NJ];
    double sums[5];
    for (int it = 0; it < 3; it++) CHECK(sfgpu_step(ctx, sp, dt, 0));
    CHECK(sfgpu_get_sums(ctx, sp, sums, &np, &n_exited, &n_slow));
    double *dep[SFGPU_NFIELDS];
    memset(dep, 0, sizeof dep);
    dep[SFGPU_F_DEN] = den;
    dep[SFGPU_F_MPC] = mpc;
    CHECK(sfgpu_get_deposit(ctx, sp, mesh, dep));
    CHECK(sfgpu_get_moments(ctx, sp, mesh, nd, NULL, NULL, NULL));
    double dsum = 0, csum = 0, ndsum = 0;
    for (int k = 0; k < NI * NJ; k++) { dsum += den[k]; csum += mpc[k]; ndsum += nd[k]; }
    if (np != N || n_exited != 0 || n_slow != 0) { fprintf(stderr, "np %lld exited %lld slow %lld\n", (long long)np, (long long)n_exited, (long long)n_slow); return 1; }
    if (csum != (double)N) { fprintf(stderr, "sum(mpc) = %.17g, expected %d\n", csum, N); return 1; }
    if (fabs(dsum - wsum) > 1e-10 * wsum || fabs(ndsum - wsum) > 1e-10 * wsum || fabs(sums[0] - wsum) > 1e-10 * wsum) {
        fprintf(stderr, "sum(Den) %.17g sum(nd) %.17g N_sum %.17g, expected %.17g\n", dsum, ndsum, sums[0], wsum);
        return 1;
    }
    /* free flight in a uniform field: u after the -0.5dt rewind and three kicks = u0 + 2.5 * (q/m) * E * dt */
    static double xo[N], yo[N], zo[N], uo[N], vo[N], wo[N], mo[N], lio[N], ljo[N], dto[N];
    static int32_t ido[N], bo[N];
    sfgpu_particles o;
    o.n = N; o.x = xo; o.y = yo; o.z = zo; o.u = uo; o.v = vo; o.w = wo; o.mpw = mo; o.li = lio; o.lj = ljo; o.dt = dto; o.id = ido; o.born_it = bo;
    CHECK(sfgpu_download(ctx, sp, mesh, 0, &o));
    const double qm = 1.602176565e-19 / (16 * 1.660538921e-27);
    for (int k = 0; k < N; k++) {
        const int id = ido[k];
        if (id < 0 || id >= N) { fprintf(stderr, "bad id %d\n", id); return 1; }
        const double want = u[id] + 2.5 * qm * 25.0 * dt;
        if (fabs(uo[k] - want) > 1e-9 * fabs(want) + 1e-9) { fprintf(stderr, "particle %d: u %.17g, expected %.17g\n", id, uo[k], want); return 1; }
        if (lio[k] != (xo[k] - x0[0]) / dh[0]) { fprintf(stderr, "particle %d: lc[0] is not XtoL(pos)\n", id); return 1; }
    }
    CHECK(sfgpu_launch_count(ctx, &nlaunch));
    CHECK(sfgpu_sync(ctx));
    sfgpu_destroy(ctx);
    printf("abi_harness OK: %d particles, 3 steps, %lld kernel launches, sum(Den) %.6f\n", N, (long long)nlaunch, dsum);
    return 0;
}
