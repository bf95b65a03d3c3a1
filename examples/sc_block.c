/*
 * examples/sc_block.c -- device-resident run of the reference's default problem from plain C, through the C ABI only.
 *
 * The reference's drivers are literal-edited main()s whose O(N^2) set-up, 32-bit CSR offsets and per-step text dumps
 * stop at ~10^5 particles.  This is the scaled stand-in: the same physics and the same
 * order of operations as src/lpmc_project.c:75-546 (3-D simple-cubic block, J2 elastoplasticity with isotropic
 * hardening + nonlocal ductile damage, top z-layer held in z, bottom z-layer loaded), every array resident in HBM,
 * set-up in O(N) on the device, compact binary snapshots instead of text dumps.
 *
 *   sc_block [n=21] [steps=3] [snapshot_every=0] [snapshot_prefix=sc_block]
 *
 * n = 21 is the default case C1 itself (9 261 particles): it prints the known answers of SURVEY section 8(c)
 * (Newton iterations 2 2 1, CG iterations 80 and 106 in load step 1).  n = 216 is the 10 077 696-particle case of
 * BASELINE.json config 5 (needs ~110 GB of HBM).  Build: see examples/Makefile (links ../lpm-c_b200/liblpmb200.so).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lpmb200.h"

#define TOLITER 1e-4 /* include/lpm.h:41 */
#define MAXITER 100  /* include/lpm.h:39 */

#define CK(call)                                                                                  \
    do {                                                                                          \
        int rc__ = (call);                                                                        \
        if (rc__ != LPMB_OK && rc__ != LPMB_ERR_NOTCONVERGED) {                                   \
            fprintf(stderr, "sc_block: %s failed (%d): %s\n", #call, rc__, lpmb_last_error());    \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 21;
    const int steps = argc > 2 ? atoi(argv[2]) : 3;
    const int snap_every = argc > 3 ? atoi(argv[3]) : 0;
    const char *prefix = argc > 4 ? argv[4] : "sc_block";
    if (n < 4 || steps < 1) {
        fprintf(stderr, "usage: sc_block [n>=4] [steps>=1] [snapshot_every] [snapshot_prefix]\n");
        return 2;
    }
    if (lpmb_device_count() < 1) {
        fprintf(stderr, "sc_block: no CUDA device (this path has no CPU fallback)\n");
        return 1;
    }
    /* material and model constants of the default driver (src/lpmc_project.c:75-260) */
    const double radius = 0.2499999944120646, h = 2.0 * radius;
    const double E0 = 146e3, mu0 = 0.3, sigmay0 = 200.0, J2_H = 38.714e3, J2_xi = 0.0;
    const double C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0), C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0),
                 C44 = E0 / 2.0 / (1.0 + mu0);
    const long long N = (long long)n * n * n;
    const int nn = 18, nconn = 61, dim = 3, plmode = 0, ntype = 4;

    double t0 = now();
    lpmb_ctx *ctx = NULL;
    CK(lpmb_create(&ctx, 0, (int)N, dim, LPMB_LATTICE_SC, nn, nconn));
    CK(lpmb_set_param(ctx, "radius", radius));
    CK(lpmb_set_param(ctx, "particle_volume", pow(2.0 * radius, 3)));       /* initialization.c:259 */
    CK(lpmb_set_param(ctx, "J2_H", J2_H));
    CK(lpmb_set_param(ctx, "J2_xi", J2_xi));
    CK(lpmb_set_param(ctx, "damage_L", 0.5));
    CK(lpmb_set_param(ctx, "damage_threshold", 0.9));
    CK(lpmb_set_param(ctx, "damagec_A", 0.0));

    /* lattice: x fastest, z slowest (initialization.c:266-284); type 1 = top z-layer, 2 = bottom z-layer,
     * 3 = particles with a complete neighbour list, 0 = the rest (lpmc_project.c:179-182) */
    double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    int *type = (int *)malloc(sizeof(int) * N);
    double *sig = (double *)malloc(sizeof(double) * N);
    for (long long i = 0; i < N; i++) {
        const int ix = (int)(i % n), iy = (int)((i / n) % n), iz = (int)(i / ((long long)n * n));
        xyz[3 * i] = -0.2 + h * ix;
        xyz[3 * i + 1] = -0.2 + h * iy;
        xyz[3 * i + 2] = -0.2 + h * iz;
        const int inner = ix > 0 && ix < n - 1 && iy > 0 && iy < n - 1 && iz > 0 && iz < n - 1;
        type[i] = iz == n - 1 ? 1 : (iz == 0 ? 2 : (inner ? 3 : 0));
        sig[i] = sigmay0;
    }
    CK(lpmb_field_set(ctx, "xyz", xyz, 3 * N));
    CK(lpmb_field_set(ctx, "xyz_initial", xyz, 3 * N));
    CK(lpmb_build_topology(ctx, 2.0 * radius, 2.0 * sqrt(2.0) * radius));  /* initialization.c:254-255 */
    CK(lpmb_field_set(ctx, "type", type, N));
    CK(lpmb_field_set(ctx, "sigmay", sig, N));
    double Ce[4 * 3];
    for (int k = 0; k < ntype; k++)
        Ce[3 * k] = C11, Ce[3 * k + 1] = C12, Ce[3 * k + 2] = C44;
    CK(lpmb_calc_kntv(ctx, Ce, ntype));
    CK(lpmb_compute_dl(ctx));
    if (N >= (1 << 18) && lpmb_matrix_enable_bricks(ctx, 1) != LPMB_OK)
        fprintf(stderr, "sc_block: brick SpMV not used (%s)\n", lpmb_last_error());
    long long nnz = 0, nblk = 0;
    CK(lpmb_csr_sizes(ctx, &nnz, &nblk));
    CK(lpmb_synchronize(ctx));
    printf("Particle number is %lld, stiffness matrix size is %lld (upper non-zeros), set-up %.2f s\n", N, nnz, now() - t0);

    /* cyclic force-controlled loading of the default driver reduced to its first branch: -2000 per step on type 2 */
    for (int step = 1; step <= steps; step++) {
        const double ts = now();
        CK(lpmb_field_copy(ctx, "xyz_temp", "xyz"));                       /* lpmc_project.c:387-389 */
        CK(lpmb_field_copy(ctx, "F_temp", "F"));
        CK(lpmb_field_copy(ctx, "Pex_temp", "Pex"));
        int newton = 0, ncg = 0, broken_total = 0, cg[MAXITER];
        CK(lpmb_fd_stiffness(ctx, 1));                                      /* :393-396 */
        CK(lpmb_apply_disp_bc(ctx, 1, 'z', 0.0));                           /* :402 */
        CK(lpmb_apply_force_bc(ctx, 2, 0.0, 0.0, -2000.0));                 /* :403 */
        CK(lpmb_bond_force(ctx, 4, 1));                                     /* :405 predictor */
        for (;;) {
            double nr = 0, nf = 0;
            CK(lpmb_update_rr(ctx, &nr, &nf));                              /* :409-414 */
            const double tol = nr > nf ? nr : nf;
            int ni = 0;
            while (nr > TOLITER * tol && ni < MAXITER) {                    /* :424-465 */
                int it = 0;
                CK(lpmb_newton_iteration(ctx, plmode, 1, 1e-8, 1e-12, (int)(3 * N), &it, &nr));
                if (ncg < MAXITER)
                    cg[ncg++] = it;
                ni++;
            }
            newton += ni;
            int broken = 0;
            CK(lpmb_update_damage(ctx, plmode, &broken, NULL, 0));          /* :469 */
            CK(lpmb_update_crack(ctx));                                     /* :470 */
            CK(lpmb_switch_state(ctx, 1));                                  /* :471 */
            broken_total += broken;
            if (broken <= 0)
                break;
            CK(lpmb_fd_stiffness(ctx, 1));                                  /* :525-541 */
        }
        CK(lpmb_synchronize(ctx));
        printf("Loading step %d has finished in %d iterations; CG iterations:", step, newton);
        for (int k = 0; k < ncg; k++)
            printf(" %d", cg[k]);
        printf("; broken bonds %d; %.3f s\n", broken_total, now() - ts);
        if (snap_every > 0 && step % snap_every == 0) {
            char path[512];
            snprintf(path, sizeof path, "%s_step%04d.lpmb", prefix, step);
            CK(lpmb_snapshot_save(ctx, path));
            printf("snapshot %s\n", path);
        }
    }
    /* mean z-displacement of the loaded layer (what result_disp.txt records for the default case) */
    double *x1 = (double *)malloc(sizeof(double) * 3 * N);
    CK(lpmb_field_get(ctx, "xyz", x1, 3 * N));
    double uz = 0;
    long long cnt = 0;
    for (long long i = 0; i < N; i++)
        if (type[i] == 2) {
            uz += x1[3 * i + 2] - xyz[3 * i + 2];
            cnt++;
        }
    printf("mean z-displacement of the loaded layer after %d steps: %.8e\n", steps, uz / (double)cnt);
    printf("kernels launched: %lld; total %.2f s\n", lpmb_launch_count(ctx), now() - t0);
    free(x1);
    free(xyz);
    free(type);
    free(sig);
    lpmb_destroy(ctx);
    return 0;
}
