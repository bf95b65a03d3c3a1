/*
 * examples/sc_block_mgpu.c -- the device-resident default problem of examples/sc_block.c on SEVERAL GPUs of one box, from
 * plain C over the C ABI only: one process per GPU, each holding a z-slab of the lattice (contiguous particle index
 * range + 4 ghost layers towards each neighbour), exactly the decomposition bench.py uses (lpm-c_b200/partition.py).
 *
 *   sc_block_mgpu <world> [n=48] [steps=1] [physics=c1|c5] [strain_step=0.005] [solver=parity|fast]
 *
 * solver fast: the opt-in preconditioned mode of the solve (param cg_precond = 1: CG preconditioned with
 * the matrix-free multigrid V-cycle, lpmb_mg.cu) -- NOT the parity path; the iteration counts printed are then PCG iterations.
 *
 * physics c1 (default): the default driver's problem (src/lpmc_project.c: E = 146e3, nu = 0.3, sigma_y = 200, force-controlled
 * -2000 per step on the bottom layer).  physics c5: BASELINE config 5 -- the material and damage law of
 * examples/CT_sc_ductile_nonlocal.c (:171,195-202,230-235: E = 115e3, nu = 0.28, sigma_y = 955, H = 2401.8, damagec_A = 400,
 * damage_L = 0.6, threshold 0.85) on the synthetic block, displacement-controlled like that example: bottom z-layer held,
 * top z-layer moved by strain_step * height every load step (yield strain 0.83 %: plastic from step 2 at the default
 * 0.5 %); every load step runs the reference's loop body (lpmc_project.c:382-546): FD tangent, BCs, predictor, Newton
 * iterations, updateDamageGeneral (nonlocal Gaussian gather over 3 * damage_L, constitutive.c:1757-1862, with its halo
 * exchange), updateCrack, switchStateV(1), re-assembly while bonds break.  The time of the damage update is printed.
 *
 * The parent starts `world` copies of itself (fork + exec, so no process inherits an initialised CUDA runtime); rank 0
 * creates the 128-byte NCCL id with lpmb_dist_unique_id and hands it to the others through a file in a private temporary
 * directory -- no MPI, no torch.  Inside a rank the call sequence is the single-GPU one (src/lpmc_project.c:382-546);
 * the halo exchange of the CG search direction, the all-reduces of its two scalars and the ghost refreshes of xyz / the
 * damage fields happen inside the same entry points once lpmb_dist_init + lpmb_dist_set_slab have been called.
 * Requirements: every rank owns at least 4 lattice layers (n >= 4 * world).  Exit code 0 = all ranks finished.
 *
 * tests/test_dist_gpu.py runs it when two devices are visible and compares the iteration counts with the single-GPU
 * known answers.
 */
#define _GNU_SOURCE
#include <math.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "lpmb200.h"

#define TOLITER 1e-4 /* include/lpm.h:41 */
#define MAXITER 100  /* include/lpm.h:39 */
#define GHOST 4      /* ghost layers per inner side: 2 with complete 2-hop stars + 2 position-only */
#define NARROW 2     /* layers exchanged per CG iteration = reach of conn */

static int g_rank = -1;

#define CK(call)                                                                                                    \
    do {                                                                                                            \
        int rc__ = (call);                                                                                          \
        if (rc__ != LPMB_OK && rc__ != LPMB_ERR_NOTCONVERGED) {                                                     \
            fprintf(stderr, "sc_block_mgpu[rank %d]: %s failed (%d): %s\n", g_rank, #call, rc__, lpmb_last_error()); \
            exit(1);                                                                                                \
        }                                                                                                           \
    } while (0)

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* owned layers [z0, z1) of rank r: the first nz % world ranks get one layer more (partition.owned_layers) */
static void owned_layers(int nz, int r, int world, int *z0, int *z1)
{
    const int base = nz / world, rem = nz % world;
    *z0 = r * base + (r < rem ? r : rem);
    *z1 = *z0 + base + (r < rem ? 1 : 0);
}

static int imin(int a, int b) { return a < b ? a : b; }

/* ghost depth of rank r towards rank-1 (lo) / rank+1 (hi) */
static void ghosts(int nz, int r, int world, int *lo, int *hi)
{
    int a, b;
    owned_layers(nz, r, world, &a, &b);
    *lo = r > 0 ? imin(GHOST, a) : 0;
    *hi = r < world - 1 ? imin(GHOST, nz - b) : 0;
}

static int run_rank(int rank, int world, int n, int steps, const char *dir, int c5, double strain_step, int fast)
{
    g_rank = rank;
    if (lpmb_device_count() < world) {
        fprintf(stderr, "sc_block_mgpu[rank %d]: %d CUDA device(s) visible, %d needed (this path has no CPU fallback)\n", rank,
                lpmb_device_count(), world);
        return 1;
    }
    /* ---- NCCL id: rank 0 creates it, the others wait for the file */
    unsigned char id[128];
    char path[600], tmp[600];
    snprintf(path, sizeof path, "%s/uid.bin", dir);
    snprintf(tmp, sizeof tmp, "%s/uid.tmp", dir);
    if (rank == 0) {
        CK(lpmb_dist_unique_id(id));
        FILE *f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, sizeof id, f) != sizeof id || fclose(f) != 0 || rename(tmp, path) != 0) {
            fprintf(stderr, "sc_block_mgpu: cannot write %s\n", path);
            return 1;
        }
    } else {
        const double t0 = now();
        FILE *f = NULL;
        while (!(f = fopen(path, "rb"))) {
            if (now() - t0 > 60.0) {
                fprintf(stderr, "sc_block_mgpu[rank %d]: no NCCL id from rank 0\n", rank);
                return 1;
            }
            usleep(20000);
        }
        if (fread(id, 1, sizeof id, f) != sizeof id) {
            fprintf(stderr, "sc_block_mgpu[rank %d]: short NCCL id\n", rank);
            return 1;
        }
        fclose(f);
    }
    /* ---- this rank's slab: layers [z0 - g_lo, z1 + g_hi) of the n^3 block, global order kept */
    int z0, z1, g_lo, g_hi, lo_of_hi, hi_of_lo, dummy;
    owned_layers(n, rank, world, &z0, &z1);
    ghosts(n, rank, world, &g_lo, &g_hi);
    /* what the neighbours expect from me = THEIR ghost depth on the side facing me */
    hi_of_lo = 0, lo_of_hi = 0;
    if (rank > 0)
        ghosts(n, rank - 1, world, &dummy, &hi_of_lo);
    if (rank < world - 1)
        ghosts(n, rank + 1, world, &lo_of_hi, &dummy);
    const long long L = (long long)n * n;                       /* particles per layer */
    const long long N = (z1 - z0 + g_lo + g_hi) * L, first = (long long)(z0 - g_lo) * L;
    const int own0 = (int)(g_lo * L), own1 = (int)((g_lo + z1 - z0) * L);

    const double radius = c5 ? 0.25 : 0.2499999944120646, h = 2.0 * radius;
    const double E0 = c5 ? 115e3 : 146e3, mu0 = c5 ? 0.28 : 0.3, sigmay0 = c5 ? 955.0 : 200.0, J2_H = c5 ? 2401.8 : 38.714e3, J2_xi = 0.0;
    const double C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0), C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0),
                 C44 = E0 / 2.0 / (1.0 + mu0);
    const int nn = 18, nconn = 61, dim = 3, plmode = 0, ntype = 4;

    const double t0 = now();
    lpmb_ctx *ctx = NULL;
    CK(lpmb_create(&ctx, rank, (int)N, dim, LPMB_LATTICE_SC, nn, nconn));
    CK(lpmb_dist_init(ctx, id, rank, world));
    CK(lpmb_dist_set_slab(ctx, own0, own1, (int)(imin(NARROW, g_lo) * L), (int)(imin(NARROW, g_hi) * L), (int)(imin(NARROW, hi_of_lo) * L),
                          (int)(imin(NARROW, lo_of_hi) * L), (int)(hi_of_lo * L), (int)(lo_of_hi * L)));
    CK(lpmb_set_param(ctx, "radius", radius));
    CK(lpmb_set_param(ctx, "particle_volume", pow(2.0 * radius, 3)));
    CK(lpmb_set_param(ctx, "J2_H", J2_H));
    CK(lpmb_set_param(ctx, "J2_xi", J2_xi));
    if (fast)
        CK(lpmb_set_param(ctx, "cg_precond", 1.0));
    CK(lpmb_set_param(ctx, "damage_L", c5 ? 0.6 : 0.5));
    CK(lpmb_set_param(ctx, "damage_threshold", c5 ? 0.85 : 0.9));
    CK(lpmb_set_param(ctx, "damagec_A", c5 ? 400.0 : 0.0));
    double *xyz = (double *)malloc(sizeof(double) * 3 * N);
    int *type = (int *)malloc(sizeof(int) * N);
    double *sig = (double *)malloc(sizeof(double) * N);
    for (long long k = 0; k < N; k++) {
        const long long i = first + k;                          /* global particle index: x fastest, z slowest */
        const int ix = (int)(i % n), iy = (int)((i / n) % n), iz = (int)(i / L);
        xyz[3 * k] = -0.2 + h * ix;
        xyz[3 * k + 1] = -0.2 + h * iy;
        xyz[3 * k + 2] = -0.2 + h * iz;
        const int inner = ix > 0 && ix < n - 1 && iy > 0 && iy < n - 1 && iz > 0 && iz < n - 1;
        type[k] = iz == n - 1 ? 1 : (iz == 0 ? 2 : (inner ? 3 : 0));
        sig[k] = sigmay0;
    }
    CK(lpmb_field_set(ctx, "xyz", xyz, 3 * N));
    CK(lpmb_field_set(ctx, "xyz_initial", xyz, 3 * N));
    CK(lpmb_build_topology(ctx, 2.0 * radius, 2.0 * sqrt(2.0) * radius));
    CK(lpmb_field_set(ctx, "type", type, N));
    CK(lpmb_field_set(ctx, "sigmay", sig, N));
    double Ce[4 * 3];
    for (int k = 0; k < ntype; k++)
        Ce[3 * k] = C11, Ce[3 * k + 1] = C12, Ce[3 * k + 2] = C44;
    CK(lpmb_calc_kntv(ctx, Ce, ntype));
    CK(lpmb_compute_dl(ctx));
    /* brick-blocked symmetric SpMV: a collective call in slab runs (every rank or none) */
    if ((long long)n * n * n >= (1 << 18) && lpmb_matrix_enable_bricks(ctx, 1) != LPMB_OK)
        fprintf(stderr, "sc_block_mgpu[rank %d]: brick SpMV not used (%s)\n", rank, lpmb_last_error());
    CK(lpmb_synchronize(ctx));
    if (rank == 0)
        printf("%d ranks, lattice %d^3 = %lld particles, %d..%d owned layers per rank, communication mode %d, physics %s, set-up %.2f s\n", world,
               n, (long long)n * n * n, n / world, (n + world - 1) / world, lpmb_dist_mode(ctx), c5 ? (fast ? "c5, fast mode (multigrid PCG)" : "c5") : (fast ? "c1, fast mode (multigrid PCG)" : "c1"), now() - t0);

    for (int step = 1; step <= steps; step++) {
        const double ts = now();
        CK(lpmb_field_copy(ctx, "xyz_temp", "xyz"));
        CK(lpmb_field_copy(ctx, "F_temp", "F"));
        CK(lpmb_field_copy(ctx, "Pex_temp", "Pex"));
        int newton = 0, ncg = 0, broken_total = 0, cg[MAXITER];
        double t_fd = 0, t_dam = 0, tq = now();
        CK(lpmb_fd_stiffness(ctx, 1));
        CK(lpmb_synchronize(ctx));
        t_fd += now() - tq;
        if (c5) {   /* displacement control (CT_sc_ductile_nonlocal.c:263-275): type 2 = bottom layer held, type 1 = top layer moved */
            CK(lpmb_apply_disp_bc(ctx, 2, 'z', 0.0));
            CK(lpmb_apply_disp_bc(ctx, 1, 'z', strain_step * h * (n - 1)));
        } else {
            CK(lpmb_apply_disp_bc(ctx, 1, 'z', 0.0));
            CK(lpmb_apply_force_bc(ctx, 2, 0.0, 0.0, -2000.0));  /* shared by the loaded layer of the WHOLE lattice */
        }
        CK(lpmb_bond_force(ctx, 4, 1));
        for (;;) {
            double nr = 0, nf = 0;
            CK(lpmb_update_rr(ctx, &nr, &nf));                   /* global norms (all-reduced) */
            const double tol = nr > nf ? nr : nf;
            int ni = 0;
            while (nr > TOLITER * tol && ni < MAXITER) {
                int it = 0;
                CK(lpmb_newton_iteration(ctx, plmode, 1, 1e-8, 1e-12, 3 * n * n * n, &it, &nr));
                if (ncg < MAXITER)
                    cg[ncg++] = it;
                ni++;
            }
            newton += ni;
            int broken = 0;
            CK(lpmb_synchronize(ctx));
            tq = now();
            CK(lpmb_update_damage(ctx, plmode, &broken, NULL, 0));   /* global count */
            CK(lpmb_synchronize(ctx));
            t_dam += now() - tq;
            CK(lpmb_update_crack(ctx));
            CK(lpmb_switch_state(ctx, 1));
            broken_total += broken;
            if (broken <= 0)
                break;
            tq = now();
            CK(lpmb_fd_stiffness(ctx, 1));
            CK(lpmb_synchronize(ctx));
            t_fd += now() - tq;
        }
        CK(lpmb_synchronize(ctx));
        if (rank == 0) {
            printf("Loading step %d has finished in %d iterations; CG iterations:", step, newton);
            for (int k = 0; k < ncg; k++)
                printf(" %d", cg[k]);
            printf("; broken bonds %d; %.3f s (FD assembly %.3f s, damage update %.3f s)\n", broken_total, now() - ts, t_fd, t_dam);
            fflush(stdout);
        }
    }
    if (c5) {   /* how far the damage law got: maximum of damage_nonlocal (slot 0) over this rank's owned particles */
        double *dn = (double *)malloc(sizeof(double) * N), *al = (double *)malloc(sizeof(double) * N);
        CK(lpmb_field_get(ctx, "damage_nonlocal0", dn, N));
        CK(lpmb_field_get(ctx, "J2_alpha0", al, N));
        double dmax = 0, amax = 0;
        long long nyield = 0;
        for (long long k = own0; k < own1; k++) {
            dmax = dn[k] > dmax ? dn[k] : dmax;
            amax = al[k] > amax ? al[k] : amax;
            nyield += al[k] > 0.0;
        }
        printf("rank %d: yielded particles %lld of %d owned, max equivalent plastic strain %.4e, max nonlocal damage %.4e\n", rank, nyield,
               own1 - own0, amax, dmax);
        fflush(stdout);
        free(dn);
        free(al);
    }
    /* the rank that owns the loaded (bottom) layer reports its mean z-displacement */
    if (rank == 0) {
        double *x1 = (double *)malloc(sizeof(double) * 3 * N);
        CK(lpmb_field_get(ctx, "xyz", x1, 3 * N));
        double uz = 0;
        long long cnt = 0;
        for (long long k = own0; k < own1; k++)
            if (type[k] == 2) {
                uz += x1[3 * k + 2] - xyz[3 * k + 2];
                cnt++;
            }
        printf("mean z-displacement of the loaded layer after %d steps: %.8e\n", steps, uz / (double)cnt);
        printf("kernels launched by rank 0: %lld; total %.2f s\n", lpmb_launch_count(ctx), now() - t0);
        free(x1);
    }
    free(xyz);
    free(type);
    free(sig);
    lpmb_destroy(ctx);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 10 && strcmp(argv[1], "--rank") == 0)   /* child: --rank r world n steps dir c5 strain_step fast */
        return run_rank(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), argv[6], atoi(argv[7]), atof(argv[8]), atoi(argv[9]));
    const int world = argc > 1 ? atoi(argv[1]) : 0;
    const int n = argc > 2 ? atoi(argv[2]) : 48;
    const int steps = argc > 3 ? atoi(argv[3]) : 1;
    const int c5 = argc > 4 && strcmp(argv[4], "c5") == 0;
    const double strain_step = argc > 5 ? atof(argv[5]) : 0.005;
    const int fast = argc > 6 && strcmp(argv[6], "fast") == 0;
    if (world < 1 || world > 16 || steps < 1 || n < 4 * world || (argc > 4 && !c5 && strcmp(argv[4], "c1") != 0) || 0) {
        fprintf(stderr, "usage: sc_block_mgpu <world 1..16> [n >= 4*world] [steps >= 1] [c1|c5] [strain_step] [parity|fast]\n");
        return 2;
    }
    char dir[] = "/tmp/lpmb_mgpu_XXXXXX";
    if (!mkdtemp(dir)) {
        perror("mkdtemp");
        return 1;
    }
    pid_t pid[16];
    char a_rank[16], a_world[16], a_n[16], a_steps[16], a_c5[4], a_strain[40], a_fast[4];
    snprintf(a_fast, sizeof a_fast, "%d", fast);
    snprintf(a_c5, sizeof a_c5, "%d", c5);
    snprintf(a_strain, sizeof a_strain, "%.17g", strain_step);
    snprintf(a_world, sizeof a_world, "%d", world);
    snprintf(a_n, sizeof a_n, "%d", n);
    snprintf(a_steps, sizeof a_steps, "%d", steps);
    for (int r = 0; r < world; r++) {
        pid[r] = fork();
        if (pid[r] < 0) {
            perror("fork");
            return 1;
        }
        if (pid[r] == 0) {
            snprintf(a_rank, sizeof a_rank, "%d", r);
            char *args[] = {argv[0], "--rank", a_rank, a_world, a_n, a_steps, dir, a_c5, a_strain, a_fast, NULL};
            execv("/proc/self/exe", args);
            perror("execv");
            _exit(127);
        }
    }
    /* a rank that dies would leave the others waiting in a collective: the first failure takes the rest down */
    int failed = 0;
    for (int left = world; left > 0; left--) {
        int st = 0;
        const pid_t p = wait(&st);
        if (p < 0)
            break;
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
            int r = 0;
            while (r < world && pid[r] != p)
                r++;
            fprintf(stderr, "sc_block_mgpu: rank %d failed (status 0x%x)\n", r, st);
            if (!failed)
                for (int q = 0; q < world; q++)
                    if (pid[q] != p)
                        kill(pid[q], SIGTERM);
            failed = 1;
        }
    }
    char path[600];
    snprintf(path, sizeof path, "%s/uid.bin", dir);
    unlink(path);
    rmdir(dir);
    return failed;
}
