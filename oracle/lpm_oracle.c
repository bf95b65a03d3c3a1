/*
 * oracle/lpm_oracle.c -- CPU restatement ("port") of LPM-C's hot path on flat arrays.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a checker (and available to bench.py's cpu_baseline leg);
 * never linked, imported or executed by the product path.  Every function cites the reference lines it
 * follows.  It is pinned (tests/test_oracle_port.py) bit-for-bit against the reference's own functions
 * running from oracle/_ref (the unmodified sources) on the committed golden case, and its CG reproduces the
 * reference's iteration counts (80 / 106 on the default case).  calcKnTv for the five lattices, crystal plasticity (plmode 1), the alternative
 * J2 laws (plmode 3 and 5), the brittle and the three remaining ductile-damage laws, the per-particle law entry points and computeStrain are restated at the end of the file as the reference's literal serial loops and
 * pinned bit-for-bit against tests/golden/sc6_j2energy.npz, sc6_j2iso.npz, sc6_damage_variants.npz, sc6_particle.npz,
 * sc6_particle2.npz (plmode 3 / 5 called per particle: the *_range entry points), fcc_cp.npz, fcc_cp_particle.npz
 * (plmode 1 called per particle with its memo: oracle_cp_particle), bcc_cp.npz, hex2d_brittle.npz, sq2d_brittle.npz and sc6_j2.npz; the
 * topology restatement also against the reference's lists at the real sizes of BASELINE configs 2-5
 * (tests/test_oracle_ref.py::test_topology_known_answers_at_the_real_config_sizes).
 *
 * Layouts are the reference's logical ones, flattened row-major: per-bond a[i*nn+j], per-particle a[i*c+k],
 * DoF vectors v[dim*i+k], Pin[3*i+k]; three-slot state as separate arrays.  Strict IEEE: compile with
 * -ffp-contract=off (oracle/Makefile).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-6     /* include/lpm.h:42 */
#define PI 3.14159265358979323846

static double dot_pairwise(int n, const double *a, const double *b);

/* ------------------------------------------------------------------ neighbor.c:9-46 (O(N^2) as there) */
int oracle_search_neighbors(int N, const double *xyz, double cutoff1, double cutoff2, int nn, int *neighbors, int *nsign, int *nb,
                            double *dist0, double *csx0, double *csy0, double *csz0)
{
    int overflow = 0;
    for (long k = 0; k < (long)N * nn; k++) {
        neighbors[k] = -1;
        nsign[k] = -1;
    }
#pragma omp parallel for schedule(static) reduction(| : overflow)
    for (int i = 0; i < N; i++) {
        int idx = 0;
        for (int j = 0; j < N; j++) {
            const double dx = xyz[3 * j] - xyz[3 * i], dy = xyz[3 * j + 1] - xyz[3 * i + 1], dz = xyz[3 * j + 2] - xyz[3 * i + 2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            int s = -1;
            if ((dis < 1.01 * cutoff1) && (j != i))
                s = 0;
            else if ((dis > 1.01 * cutoff1) && (dis < 1.01 * cutoff2))
                s = 1;
            if (s < 0)
                continue;
            if (idx >= nn) {
                overflow = 1;
                continue;
            }
            const long e = (long)i * nn + idx;
            csx0[e] = (xyz[3 * i] - xyz[3 * j]) / dis;
            csy0[e] = (xyz[3 * i + 1] - xyz[3 * j + 1]) / dis;
            csz0[e] = (xyz[3 * i + 2] - xyz[3 * j + 2]) / dis;
            dist0[e] = dis;
            nsign[e] = s;
            neighbors[e] = j;
            idx++;
        }
        nb[i] = idx;
    }
    return overflow;
}

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

/* ------------------------------------------------------------------ neighbor.c:49-130 */
/* conn[i] = sorted unique( {j} U N1(j) for first-shell j, {j} U N2(j) for second-shell j ); K_pointer (64-bit) */
int oracle_afem_conn(int N, int nn, int nconn, int dim, const int *neighbors, const int *nsign, const int *nb, int *conn, int *nb_conn,
                     long long *kp0, long long *kp1)
{
    int overflow = 0;
    int *tmp = (int *)malloc(sizeof(int) * (size_t)nn * (nn + 1));
    for (int i = 0; i < N; i++) {
        int n = 0;
        for (int j = 0; j < nb[i]; j++) {
            const int nj = neighbors[(long)i * nn + j], s = nsign[(long)i * nn + j];
            tmp[n++] = nj;
            for (int m = 0; m < nb[nj]; m++)
                if (nsign[(long)nj * nn + m] == s)
                    tmp[n++] = neighbors[(long)nj * nn + m];
        }
        qsort(tmp, n, sizeof(int), cmp_int);
        int u = 0;
        for (int k = 0; k < n; k++)
            if (k == 0 || tmp[k] != tmp[k - 1]) {
                if (u < nconn)
                    conn[(long)i * nconn + u] = tmp[k];
                else
                    overflow = 1;
                u++;
            }
        nb_conn[i] = u < nconn ? u : nconn;
        for (int k = nb_conn[i]; k < nconn; k++)
            conn[(long)i * nconn + k] = -1;
    }
    free(tmp);
    kp1[0] = 0;
    for (int i = 0; i < N; i++) {
        int ge = 0;
        for (int j = 0; j < nb_conn[i]; j++)
            if (conn[(long)i * nconn + j] >= i)
                ge++;
        kp0[i] = ge;
        kp1[i + 1] = kp1[i] + (long long)dim * dim * ge - (dim == 3 ? 3 : 1);
    }
    kp0[N] = 0;
    return overflow;
}

/* ------------------------------------------------------------------ geometry pass G(i; dLp*) */
/* constitutive.c:241-260 / 495-515 / 625-645 (apply_broken=1) and lpm_basic.c:252-291 (apply_broken=0, writes distance) */
void oracle_geometry(int N, int nn, const double *xyz, const int *neighbors, const int *nsign, const int *nbi, const double *L0,
                     const double *dLp, const double *broken, const double *Tv, int apply_broken, double *dL, double *csx, double *csy,
                     double *csz, double *dLt, double *TdLt, double *distance)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        dLt[2 * i] = dLt[2 * i + 1] = TdLt[2 * i] = TdLt[2 * i + 1] = 0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e];
            const double dx = xyz[3 * i] - xyz[3 * nj], dy = xyz[3 * i + 1] - xyz[3 * nj + 1], dz = xyz[3 * i + 2] - xyz[3 * nj + 2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            if (distance)
                distance[e] = dis;
            dL[e] = dis - L0[e];
            dL[e] -= dLp[e];
            if (apply_broken)
                dL[e] *= broken[e];
            dLt[2 * i + nsign[e]] += dL[e];
            TdLt[2 * i + nsign[e]] += Tv[e] * dL[e];
            csx[e] = dx / dis;
            csy[e] = dy / dis;
            csz[e] = dz / dis;
        }
    }
}

/* owner pass: law 6 = elastic (constitutive.c:264-279), law 0 = J2 average stretch (constitutive.c:648-667) */
void oracle_force(int N, int nn, int law, const int *neighbors, const int *nsign, const int *nbi, const double *Kn, const double *Tv,
                  const double *scale, const double *dL, const double *dLt, const double *TdLt, const double *csx, const double *csy,
                  const double *csz, double *dL_ave, double *F, double *Pin)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], s = nsign[e];
            double stretch = dL[e];
            if (law == 0) {
                for (int jj = 0; jj < nn; jj++)
                    if (neighbors[(long)nj * nn + jj] == i)
                        dL_ave[e] = 0.5 * (dL[e] + dL[(long)nj * nn + jj]);
                stretch = dL_ave[e];
            }
            F[e] = 2.0 * Kn[e] * stretch + 0.5 * (TdLt[2 * i + s] + TdLt[2 * nj + s]) + 0.5 * Tv[e] * (dLt[2 * i + s] + dLt[2 * nj + s]);
            F[e] *= scale[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
    }
}

/* ------------------------------------------------------------------ constitutive.c:167-225 (predictor, plmode 4) */
void oracle_predictor(int N, int nn, const double *xyz, const double *xyz_temp, const int *neighbors, const int *nsign, const int *nbi,
                      const double *Kn, const double *Tv, const double *broken, const double *F_temp, const double *csx, const double *csy,
                      const double *csz, double *ddL, double *ddLt, double *TddLt, double *F, double *Pin)
{
    for (int i = 0; i < N; i++) {
        ddLt[2 * i] = ddLt[2 * i + 1] = TddLt[2 * i] = TddLt[2 * i + 1] = 0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e];
            const double ax = xyz_temp[3 * i] - xyz_temp[3 * nj], ay = xyz_temp[3 * i + 1] - xyz_temp[3 * nj + 1], az = xyz_temp[3 * i + 2] - xyz_temp[3 * nj + 2];
            const double dis0 = sqrt(ax * ax + ay * ay + az * az);
            const double bx = xyz[3 * i] - xyz[3 * nj], by = xyz[3 * i + 1] - xyz[3 * nj + 1], bz = xyz[3 * i + 2] - xyz[3 * nj + 2];
            const double dis1 = sqrt(bx * bx + by * by + bz * bz);
            ddL[e] = broken[e] * (dis1 - dis0);
            ddLt[2 * i + nsign[e]] += ddL[e];
            TddLt[2 * i + nsign[e]] += Tv[e] * ddL[e];
        }
    }
    for (int i = 0; i < N; i++) {
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], s = nsign[e];
            F[e] = F_temp[e] + 2.0 * Kn[e] * ddL[e] + 0.5 * (TddLt[2 * i + s] + TddLt[2 * nj + s]) + 0.5 * Tv[e] * (ddLt[2 * i + s] + ddLt[2 * nj + s]);
            F[e] *= broken[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
    }
}

/* opposite-bond factor: lpm_basic.c:72-90, constitutive.c:541-559 */
static double opp_flag(int i, int j, int nn, int nneighbors, const int *nb, const int *nbi, const double *cx0, const double *cy0,
                       const double *cz0, const double *broken)
{
    if (nb[i] == nneighbors)
        return 0.5;
    for (int m = 0; m < nbi[i]; m++) {
        const long a = (long)i * nn + m, b = (long)i * nn + j;
        if (fabs(cx0[a] + cx0[b]) < EPS && fabs(cy0[a] + cy0[b]) < EPS && fabs(cz0[a] + cz0[b]) < EPS)
            return broken[a] <= EPS ? 1.0 : 0.5;
    }
    return 1.0;
}

/* ------------------------------------------------------------------ constitutive.c:518-622, 669-675 (J2 return map, once per particle) */
void oracle_j2_return_map(int N, int nn, double V, double J2_H, double J2_xi, const double *Ce, const int *type, const double *sigmay,
                          const int *nsign, const int *nb, const int *nbi, const double *Kn, const double *Tv, const double *w,
                          const double *broken, const double *L0, const double *cx0, const double *cy0, const double *cz0, const double *dL,
                          const double *dLt, const double *TdLt, const double *csx, const double *csy, const double *csz, const double *dLp0,
                          const double *beta0, const double *alpha0, double *dLp2, double *beta2, double *alpha2, double *ddLp,
                          double *dlambda, int *pl_flag)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        double st[6] = {0}, dpl[6] = {0};
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            double Fij = 2.0 * Kn[e] * dL[e] + TdLt[2 * i + nsign[e]] + Tv[e] * dLt[2 * i + nsign[e]];
            Fij *= w[e];
            const double of = opp_flag(i, j, nn, nn, nb, nbi, cx0, cy0, cz0, broken);
            st[0] += of / V * L0[e] * Fij * csx[e] * csx[e];
            st[1] += of / V * L0[e] * Fij * csy[e] * csy[e];
            st[2] += of / V * L0[e] * Fij * csz[e] * csz[e];
            st[3] += of / V * L0[e] * Fij * csy[e] * csz[e];
            st[4] += of / V * L0[e] * Fij * csx[e] * csz[e];
            st[5] += of / V * L0[e] * Fij * csx[e] * csy[e];
        }
        const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
        for (int q = 0; q < 3; q++)
            st[q] -= temp;
        double beta[6];
        for (int q = 0; q < 6; q++) {
            beta[q] = beta0[6 * i + q];
            st[q] -= beta[q];
        }
        double seq = 0.0;
        for (int q = 0; q < 6; q++)
            seq += (q < 3 ? 1.0 : 2.0) * st[q] * st[q];
        seq = sqrt(3.0 / 2.0 * seq);
        double alpha = alpha0[i], dl = 0.0;
        const double yf = seq - (sigmay[i] + (1.0 - J2_xi) * J2_H * alpha);
        if (yf > 0.0) {
            pl_flag[i] = 1;
            dl = yf / (3 * Ce[3 * type[i] + 2] + J2_H);
        }
        alpha += dl;
        for (int q = 0; q < 6; q++)
            if (fabs(seq) > EPS) {
                dpl[q] = dl * 1.5 * st[q] / seq;
                beta[q] += 2. / 3. * J2_xi * J2_H * dpl[q];
            }
        for (int j = 0; j < nn; j++) {
            const long e = (long)i * nn + j;
            double xd = dLp0[e];
            if (j < nbi[i]) {
                ddLp[e] = L0[e] * (dpl[0] * csx[e] * csx[e] + dpl[1] * csy[e] * csy[e] + dpl[2] * csz[e] * csz[e] + 2 * dpl[3] * csy[e] * csz[e] +
                                   2 * dpl[4] * csx[e] * csz[e] + 2 * dpl[5] * csx[e] * csy[e]);
                ddLp[e] *= broken[e];
                xd += ddLp[e];
            }
            dLp2[e] = broken[e] * xd;
        }
        for (int q = 0; q < 6; q++)
            beta2[6 * i + q] = beta[q];
        alpha2[i] = alpha;
        dlambda[i] = dl;
    }
}

/* ------------------------------------------------------------------ lpm_basic.c:53-125 */
void oracle_stress(int N, int nn, double V, const int *nb, const int *nbi, const double *L0, const double *cx0, const double *cy0,
                   const double *cz0, const double *broken, const double *F, const double *csx, const double *csy, const double *csz,
                   double *stress, double *seq_out, double *sm_out, double *triax, double *bond_stress)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        double *st = stress + 6 * (long)i;
        memset(st, 0, 6 * sizeof(double));
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const double of = opp_flag(i, j, nn, nn, nb, nbi, cx0, cy0, cz0, broken);
            st[0] += of / V * L0[e] * F[e] * csx[e] * csx[e];
            st[1] += of / V * L0[e] * F[e] * csy[e] * csy[e];
            st[2] += of / V * L0[e] * F[e] * csz[e] * csz[e];
            st[3] += of / V * L0[e] * F[e] * csy[e] * csz[e];
            st[4] += of / V * L0[e] * F[e] * csx[e] * csz[e];
            st[5] += of / V * L0[e] * F[e] * csx[e] * csy[e];
        }
        double seq = 0.0;
        for (int q = 0; q < 6; q++)
            seq += (q < 3 ? 1.0 : 2.0) * st[q] * st[q];
        seq_out[i] = sqrt(3.0 / 2.0 * seq);
        sm_out[i] = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
        if (seq_out[i] > EPS)
            triax[i] = sm_out[i] / seq_out[i];
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            bond_stress[e] = (st[0] * csx[e] * csx[e] + st[1] * csy[e] * csy[e] + st[2] * csz[e] * csz[e] + 2 * st[3] * csy[e] * csz[e] +
                              2 * st[4] * csx[e] * csz[e] + 2 * st[5] * csx[e] * csy[e]);
        }
    }
}

/* ------------------------------------------------------------------ stiffness.c:519-534 */
int oracle_update_rr(int N, int dim, const int *bc, const double *Pex, const double *Pin, double *residual, double *reaction)
{
    int ii = 0;
    for (int i = 0; i < N; i++)
        for (int k = 0; k < dim; k++) {
            residual[dim * i + k] = bc[dim * i + k] * (Pex[dim * i + k] - Pin[3 * i + k]);
            if (bc[dim * i + k] == 0)
                reaction[ii++] = Pin[3 * i + k];
        }
    return ii;
}

/* ------------------------------------------------------------------ stiffness.c:271-516: FD tangent, brute force as there */
/* elastic internal force of particle ii with the CURRENT xyz (constitutive.c:228-283); scratch holds the per-particle sums */
static void elastic_pin(int ii, int nn, const double *xyz, const int *neighbors, const int *nsign, const int *nbi, const double *L0,
                        const double *dLp0, const double *broken, const double *Kn, const double *Tv, double *dL, double *csx, double *csy,
                        double *csz, double *dLt, double *TdLt, double *Fout, double *pin)
{
    for (int k = -1; k < nbi[ii]; k++) {
        int i = ii;
        if (k >= 0) {
            if (!(broken[(long)ii * nn + k] > EPS))
                continue;
            i = neighbors[(long)ii * nn + k];
        }
        dLt[2 * i] = dLt[2 * i + 1] = TdLt[2 * i] = TdLt[2 * i + 1] = 0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e];
            const double dx = xyz[3 * i] - xyz[3 * nj], dy = xyz[3 * i + 1] - xyz[3 * nj + 1], dz = xyz[3 * i + 2] - xyz[3 * nj + 2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            dL[e] = dis - L0[e];
            dL[e] -= dLp0[e];
            dL[e] *= broken[e];
            dLt[2 * i + nsign[e]] += dL[e];
            TdLt[2 * i + nsign[e]] += Tv[e] * dL[e];
            csx[e] = dx / dis;
            csy[e] = dy / dis;
            csz[e] = dz / dis;
        }
    }
    pin[0] = pin[1] = pin[2] = 0.0;
    for (int j = 0; j < nbi[ii]; j++) {
        const long e = (long)ii * nn + j;
        const int nj = neighbors[e], s = nsign[e];
        double f = 2.0 * Kn[e] * dL[e] + 0.5 * (TdLt[2 * ii + s] + TdLt[2 * nj + s]) + 0.5 * Tv[e] * (dLt[2 * ii + s] + dLt[2 * nj + s]);
        f *= broken[e];
        Fout[e] = f;
        pin[0] += csx[e] * f;
        pin[1] += csy[e] * f;
        pin[2] += csz[e] * f;
    }
}

/* serial, single pass over particles exactly like the reference run with one thread; leaves the same side effects
 * in dL, cs*, dLt, TdLt, F, Pin.  K / JK / IK are the 1-based symmetric-upper CSR. */
void oracle_fd_stiffness(int N, int nn, int nconn, int dim, double radius, double *xyz, const int *neighbors, const int *nsign, const int *nbi,
                         const double *L0, const double *dLp0, const double *broken, const double *Kn, const double *Tv, const int *conn,
                         const int *nb_conn, const long long *kp0, const long long *kp1, double *K, int *JK, int *IK, double *dL, double *csx,
                         double *csy, double *csz, double *dLt, double *TdLt, double *F, double *Pin)
{
    const int D = dim;
    memset(K, 0, sizeof(double) * (size_t)kp1[N]);
    double *Kl = (double *)malloc(sizeof(double) * 3 * nconn * 3);
    for (int i = 0; i < N; i++) {
        const int nc = nb_conn[i];
        double base[3], pin[3];
        elastic_pin(i, nn, xyz, neighbors, nsign, nbi, L0, dLp0, broken, Kn, Tv, dL, csx, csy, csz, dLt, TdLt, F, base);
        memcpy(Pin + 3 * i, base, sizeof base);
        for (int jID = 0; jID < nc; jID++)
            for (int r = 0; r < D; r++) {
                const int c = conn[(long)i * nconn + jID];
                const double xt = xyz[3 * c + r];
                xyz[3 * c + r] = xt + EPS * radius;
                elastic_pin(i, nn, xyz, neighbors, nsign, nbi, L0, dLp0, broken, Kn, Tv, dL, csx, csy, csz, dLt, TdLt, F, pin);
                memcpy(Pin + 3 * i, pin, sizeof pin);
                xyz[3 * c + r] = xt;
                for (int s = 0; s < D; s++)
                    Kl[(r * nc + jID) * 3 + s] = (pin[s] - base[s]) / EPS / radius;
            }
        const long long P = kp1[i];
        const int K0 = (int)kp0[i];
#define ROWSTART(PP, r, KK) ((PP) + (long long)(r) * D * (KK) - (long long)((r) * ((r)-1) / 2))
        int num1 = 0;
        for (int jID = 0; jID < nc; jID++) {
            const int jj = conn[(long)i * nconn + jID];
            double kll[3][3];
            for (int r = 0; r < D; r++)
                for (int s = 0; s < D; s++)
                    kll[r][s] = Kl[(s * nc + jID) * 3 + r];
            if (jj == i) {
                for (int r = 0; r < D; r++)
                    for (int s = r; s < D; s++) {
                        K[ROWSTART(P, r, K0) + (s - r)] += kll[r][s];
                        JK[ROWSTART(P, r, K0) + (s - r)] = D * jj + s + 1;
                    }
            } else if (jj > i) {
                num1++;
                for (int r = 0; r < D; r++)
                    for (int s = 0; s < D; s++) {
                        K[ROWSTART(P, r, K0) + (long long)num1 * D - r + s] += 0.5 * kll[r][s];
                        JK[ROWSTART(P, r, K0) + (long long)num1 * D - r + s] = D * jj + s + 1;
                    }
            } else {
                int num2 = 0;
                const long long Pj = kp1[jj];
                const int K0j = (int)kp0[jj];
                for (int k = 0; k < nb_conn[jj]; k++) {
                    if (conn[(long)jj * nconn + k] <= jj)
                        continue;
                    num2++;
                    if (conn[(long)jj * nconn + k] == i)
                        for (int r = 0; r < D; r++)
                            for (int s = 0; s < D; s++)
                                K[ROWSTART(Pj, r, K0j) + (long long)num2 * D - r + s] += 0.5 * kll[s][r];
                }
            }
        }
        for (int r = 0; r < D; r++)
            IK[D * i + r] = (int)(ROWSTART(P, r, K0) + 1);
    }
    IK[D * N] = (int)(kp1[N] + 1);
    free(Kl);
}

/* ------------------------------------------------------------------ boundary.c:159-281 (and 2-D twin :72-157) */
void oracle_bc_stiffness_update(int N, int nconn, int dim, const int *bc, const int *fix, const int *conn, const int *nb_conn,
                                const long long *kp0, const long long *kp1, double *K, double *residual)
{
    const int D = dim;
    /* norm_diag = cblas_dnrm2(diag) (boundary.c:168-176); the open shim sums pairwise above 4096 entries */
    double *diag = (double *)malloc(sizeof(double) * (size_t)N * D);
    for (int i = 0; i < N; i++)
        for (int r = 0; r < D; r++)
            diag[D * i + r] = K[ROWSTART(kp1[i], r, (int)kp0[i])];
    double s2 = 0.0;
    if (N * D > 4096)
        s2 = dot_pairwise(N * D, diag, diag);
    else
        for (int k = 0; k < N * D; k++)
            s2 += diag[k] * diag[k];
    free(diag);
    const double norm_diag = sqrt(s2);
    for (int i = 0; i < N; i++)
        for (int a = 0; a < D; a++) {
            if (!(bc[D * i + a] == 0 || fix[D * i + a] == 0))
                continue;
            const long long P = kp1[i];
            const int K0 = (int)kp0[i];
            /* own block rows: column a of rows r<a, the whole row a */
            for (int r = 0; r < a; r++)
                K[ROWSTART(P, r, K0) + (a - r)] = 0.0;
            const long long rs = ROWSTART(P, a, K0), re = (a + 1 < D) ? ROWSTART(P, a + 1, K0) : kp1[i + 1];
            for (long long kk = rs; kk < re; kk++)
                K[kk] = 0.0;
            K[rs] = norm_diag;
            /* column a of particle i inside the rows of lower-index conn particles */
            for (int j = 0; j < nb_conn[i]; j++) {
                const int cj = conn[(long)i * nconn + j];
                if (cj >= i)
                    continue;
                int num2 = 0;
                for (int k = 0; k < nb_conn[cj]; k++) {
                    if (conn[(long)cj * nconn + k] <= cj)
                        continue;
                    num2++;
                    if (conn[(long)cj * nconn + k] == i)
                        for (int r = 0; r < D; r++)
                            K[ROWSTART(kp1[cj], r, (int)kp0[cj]) + (long long)num2 * D - r + a] = 0.0;
                }
            }
            residual[D * i + a] = 0.0;
        }
}

/* ------------------------------------------------------------------ solver.c:188-270 with the shim's documented RCI-CG */
static double dot_pairwise(int n, const double *a, const double *b)
{
    if (n <= 32) {
        double s = 0.0;
        for (int i = 0; i < n; i++)
            s += a[i] * b[i];
        return s;
    }
    const int h = (n / 2 + 31) & ~31;
    return dot_pairwise(h, a, b) + dot_pairwise(n - h, a + h, b + h);
}

static void spmv_sym_upper(int n, const int *IK, const int *JK, const double *K, const double *x, double *y)
{
    memset(y, 0, sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = IK[i] - 1; k < IK[i + 1] - 1; k++) {
            const int j = JK[k] - 1;
            s += K[k] * x[j];
            if (j != i)
                y[j] += 1.0 * (K[k] * x[i]);
        }
        y[i] += 1.0 * s;
    }
}

/* x0 = 0; stop when ||r||^2 <= rel*||r0||^2 + abs (squared norms) or maxit.  Returns the iteration count. */
int oracle_cg(int n, const int *IK, const int *JK, const double *K, const double *b, double *x, double rel, double abs_tol, int maxit)
{
    double *p = (double *)malloc(sizeof(double) * n * 3), *ap = p + n, *r = p + 2 * n;
    memset(x, 0, sizeof(double) * n);
    memcpy(r, b, sizeof(double) * n);
    memcpy(p, b, sizeof(double) * n);
    double rr = dot_pairwise(n, r, r);
    const double thresh = rel * rr + abs_tol;
    int it = 0;
    if (rr > thresh)
        while (it < maxit) {
            spmv_sym_upper(n, IK, JK, K, p, ap);
            const double alpha = rr / dot_pairwise(n, p, ap);
            for (int i = 0; i < n; i++)
                x[i] += alpha * p[i];
            for (int i = 0; i < n; i++)
                r[i] += -alpha * ap[i];
            const double rr_new = dot_pairwise(n, r, r);
            it++;
            if (rr_new <= thresh)
                break;
            const double beta = rr_new / rr;
            for (int i = 0; i < n; i++)
                p[i] = r[i] + beta * p[i];
            rr = rr_new;
        }
    free(p);
    return it;
}

/* ------------------------------------------------------------------ constitutive.c:1757-1862 (O(N^2) as there) */
int oracle_damage_nonlocal(int N, int nn, double L, double thr, double Ac, double V, const double *xyz0, const int *neighbors, const int *nbi,
                           const double *dlambda, const double *triax, double *Dn, double *broken, double *dD0, double *w)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        if (Dn[i] > thr) {
            if (Dn[i] > 1.0)
                Dn[i] = 1.0;
            continue;
        }
        double Ddot = 0, A = 0;
        for (int j = 0; j < N; j++) {
            const double dx = xyz0[3 * j] - xyz0[3 * i], dy = xyz0[3 * j + 1] - xyz0[3 * i + 1], dz = xyz0[3 * j + 2] - xyz0[3 * i + 2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            if (dis < 3 * L) {
                double DdotLocal = 0;
                const double f = (1.0 + Ac * triax[j]);
                if (f > 0.0)
                    DdotLocal = dlambda[j] * (1.0 + Ac * triax[j]);
                const double phi = 1.0 / L / sqrt(2 * PI) * exp(-0.5 * dis * dis / L / L);
                Ddot += DdotLocal * phi * V;
                A += phi * V;
            }
        }
        if (Ddot > 0.0)
            Dn[i] += 1.0 / A * Ddot;
    }
    int k = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (Dn[i] > thr || Dn[neighbors[e]] > thr)
                if (fabs(broken[e]) > EPS) {
                    broken[e] = 0.0;
                    dD0[e] = 1.0;
                    k++;
                }
        }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (fabs(broken[e]) > EPS)
                dD0[e] = Dn[i] < Dn[neighbors[e]] ? Dn[neighbors[e]] : Dn[i];
            w[e] = 1.0 - dD0[e];
        }
    return k;
}

/* ------------------------------------------------------------------ constitutive.c:1399-1434 */
void oracle_update_crack(int N, int nn, int dim, const int *nbi, const double *broken, const double *w, const double *csx, const double *csy,
                         const double *csz, double *F, double *Pin, int *nb, double *damage_visual, int *fix_index)
{
    for (int i = 0; i < N; i++) {
        nb[i] = nbi[i];
        damage_visual[i] = 0.0;
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (broken[e] <= EPS)
                nb[i] -= 1;
            damage_visual[i] += broken[e];
            F[e] *= w[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
        if (nb[i] < 1)
            for (int k = 0; k < dim; k++)
                fix_index[dim * i + k] = 0;
        damage_visual[i] = 1 - damage_visual[i] / nbi[i];
    }
}

/* ================================================================== rows a8 / f4 of SURVEY section 8 ============
 * The alternative J2 laws are restated LITERALLY as the reference's serial particle loop (they are order
 * dependent: plmode 5 updates slot [0] in place, plmode 3 reads stale sums across broken bonds), on flat arrays. */

/* geometry of particle i with an explicit plastic-stretch row, constitutive.c:318-334 / 703-721 (cs optional) */
static void geom_row(int i, int nn, const double *xyz, const int *neighbors, const int *nsign, const int *nbi, const double *L0,
                     const double *dlp_row, const double *broken, const double *Tv, double *dL, double *dLt, double *TdLt, double *csx,
                     double *csy, double *csz)
{
    dLt[2 * i] = dLt[2 * i + 1] = TdLt[2 * i] = TdLt[2 * i + 1] = 0;
    for (int j = 0; j < nbi[i]; j++) {
        const long e = (long)i * nn + j;
        const int nj = neighbors[e];
        const double dx = xyz[3 * i] - xyz[3 * nj], dy = xyz[3 * i + 1] - xyz[3 * nj + 1], dz = xyz[3 * i + 2] - xyz[3 * nj + 2];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        dL[e] = dis - L0[e];
        dL[e] -= dlp_row[j];
        dL[e] *= broken[e];
        dLt[2 * i + nsign[e]] += dL[e];
        TdLt[2 * i + nsign[e]] += Tv[e] * dL[e];
        if (csx) {
            csx[e] = dx / dis;
            csy[e] = dy / dis;
            csz[e] = dz / dis;
        }
    }
}

/* list of ii and its intact neighbours, constitutive.c:289-295 / 692-699 */
static int star_list(int ii, int nn, const int *neighbors, const double *broken, const int *nb, int *out)
{
    int s = 0;
    out[s++] = ii;
    for (int k = 0; k < nn; k++)
        if (broken[(long)ii * nn + k] > EPS && neighbors[(long)ii * nn + k] != -1 && s < nb[ii] + 1)
            out[s++] = neighbors[(long)ii * nn + k];
    while (s < nb[ii] + 1)
        out[s++] = ii; /* allocInt1D(nb+1, ii) pre-fills with ii */
    return nb[ii] + 1;
}

/* computeBondForceJ2energyReturnMap for all particles in order, constitutive.c:286-463 (plmode 3) */
/* calls ii = ii0 .. ii1-1 of the serial loop; (0, N) = the dispatcher's loop, (ii, ii+1) = the entry point called on its own */
void oracle_j2_energy_force_range(int ii0, int ii1, int N, int nn, double V, double radius, double J2_H, double J2_xi, int load_indicator,
                                  const double *Ce, const int *type, const double *sigmay, const double *xyz, const int *neighbors,
                                  const int *nsign, const int *nbi, const int *nb, const double *L0, const double *Kn, const double *Tv,
                                  const double *broken, const double *dD0, const double *dLp0, const double *beq0, const double *alpha0,
                                  double *dLp2, double *beq2, double *alpha2, double *dlambda_out, int *pl_flag, double *ddLp, double *dL,
                                  double *dL_ave, double *dLt, double *TdLt, double *csx, double *csy, double *csz, double *F, double *Pin)
{
    (void)N;
    int *tmp = (int *)malloc(sizeof(int) * (nn + 1));
    double *xdLp = (double *)malloc(sizeof(double) * (nn + 1) * nn), *xa = (double *)malloc(sizeof(double) * (nn + 1));
    double *xb = (double *)malloc(sizeof(double) * (nn + 1)), *xl = (double *)malloc(sizeof(double) * (nn + 1));
    for (int ii = ii0; ii < ii1; ii++) {
        const int cnt = star_list(ii, nn, neighbors, broken, nb, tmp);
        for (int k = 0; k < cnt; k++) {
            const int i = tmp[k];
            for (int j = 0; j < nn; j++)
                xdLp[k * nn + j] = dLp0[(long)i * nn + j];
            xb[k] = beq0[i];
            xa[k] = alpha0[i];
            xl[k] = 0.0;
        }
        for (int k = 0; k < cnt; k++)
            geom_row(tmp[k], nn, xyz, neighbors, nsign, nbi, L0, xdLp + k * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
        for (int k = 0; k < cnt; k++) {
            const int i = tmp[k];
            int nb1 = 0;
            for (int j = 0; j < nbi[i]; j++) /* countNEqual(neighbors1[i], nneighbors1, -1) */
                nb1 += nsign[(long)i * nn + j] == 0;
            const int nb2 = nb[i] - nb1;
            double U_d = 0.0, J2_k = 0.0;
            const double J2_V = V * nb[i] / nn;
            for (int j = 0; j < nb[i]; j++) {
                const long e = (long)i * nn + j;
                if (nsign[e] == 0) {
                    U_d += 0.5 * Kn[e] * (dL[e] - dLt[2 * i] / nb1) * (dL[e] - dLt[2 * i] / nb1);
                    J2_k = Kn[e];
                } else if (nsign[e] == 1) {
                    U_d += 0.5 * Kn[e] * (dL[e] - dLt[2 * i + 1] / nb2) * (dL[e] - dLt[2 * i + 1] / nb2);
                }
            }
            const double c44 = Ce[3 * type[i] + 2];
            const double J2_sigma = sqrt(6.0 * c44 * U_d / J2_V);
            double dlambda = 0.0;
            double yield_func = fabs(load_indicator * J2_sigma - beq0[i]) - (sigmay[i] + (1.0 - J2_xi) * J2_H * alpha0[i]);
            if (yield_func > 0.0) {
                pl_flag[i] = 1;
                double a = 0.0, b = 1.0, ya = yield_func;
                while ((b - a) > 1e-4) {
                    dlambda = (a + b) / 2.0;
                    yield_func = fabs(load_indicator * J2_sigma / (1.0 + 0.5 * dlambda) -
                                      (beq0[i] + load_indicator * J2_xi * J2_H * dlambda * J2_sigma * sqrt(radius / J2_k / c44) / 6.0 / (1.0 + 0.5 * dlambda))) -
                                 (sigmay[i] + (1.0 - J2_xi) * J2_H * alpha0[i] +
                                  (1.0 - J2_xi) * J2_H * dlambda * J2_sigma * sqrt(radius / J2_k / c44) / 6.0 / (1.0 + 0.5 * dlambda));
                    if (yield_func * ya < 0.0) {
                        b = dlambda;
                    } else {
                        a = dlambda;
                        ya = yield_func;
                    }
                }
            }
            xl[k] = dlambda;
            xa[k] += dlambda / (1.0 + 0.5 * dlambda) * sqrt(radius / 6.0 / J2_k * U_d / J2_V);
            xb[k] = (xb[k] + load_indicator * J2_xi * J2_H * dlambda / (1.0 + 0.5 * dlambda) * sqrt(radius / 6.0 / J2_k * U_d / J2_V));
            double f_d = 0.0;
            for (int j = 0; j < nb[i]; j++) {
                const long e = (long)i * nn + j;
                if (nsign[e] == 0)
                    f_d = 2.0 * Kn[e] / (1 + dlambda / 2.0) * (dL[e] - dLt[2 * i] / nb1);
                if (nsign[e] == 1)
                    f_d = 2.0 * Kn[e] / (1 + dlambda / 2.0) * (dL[e] - dLt[2 * i + 1] / nb2);
                ddLp[e] = dlambda * f_d / (4.0 * Kn[e]);
                ddLp[e] *= broken[e];
                xdLp[k * nn + j] += ddLp[e];
            }
        }
        for (int k = 0; k < cnt; k++)
            geom_row(tmp[k], nn, xyz, neighbors, nsign, nbi, L0, xdLp + k * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
        const int i = ii;
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], s = nsign[e];
            for (int jj = 0; jj < nn; jj++)
                if (neighbors[(long)nj * nn + jj] == i)
                    dL_ave[e] = 0.5 * (dL[e] + dL[(long)nj * nn + jj]);
            F[e] = 2.0 * Kn[e] * dL_ave[e] + 0.5 * (TdLt[2 * i + s] + TdLt[2 * nj + s]) + 0.5 * Tv[e] * (dLt[2 * i + s] + dLt[2 * nj + s]);
            F[e] *= (1.0 - dD0[e]);
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
        for (int j = 0; j < nn; j++)
            dLp2[(long)i * nn + j] = broken[(long)i * nn + j] * xdLp[j];
        beq2[i] = xb[0];
        alpha2[i] = xa[0];
        dlambda_out[i] = xl[0];
    }
    free(tmp);
    free(xdLp);
    free(xa);
    free(xb);
    free(xl);
}

void oracle_j2_energy_force(int N, int nn, double V, double radius, double J2_H, double J2_xi, int load_indicator, const double *Ce,
                            const int *type, const double *sigmay, const double *xyz, const int *neighbors, const int *nsign, const int *nbi,
                            const int *nb, const double *L0, const double *Kn, const double *Tv, const double *broken, const double *dD0,
                            const double *dLp0, const double *beq0, const double *alpha0, double *dLp2, double *beq2, double *alpha2,
                            double *dlambda_out, int *pl_flag, double *ddLp, double *dL, double *dL_ave, double *dLt, double *TdLt,
                            double *csx, double *csy, double *csz, double *F, double *Pin)
{
    oracle_j2_energy_force_range(0, N, N, nn, V, radius, J2_H, J2_xi, load_indicator, Ce, type, sigmay, xyz, neighbors, nsign, nbi, nb, L0, Kn, Tv,
                                 broken, dD0, dLp0, beq0, alpha0, dLp2, beq2, alpha2, dlambda_out, pl_flag, ddLp, dL, dL_ave, dLt, TdLt, csx, csy,
                                 csz, F, Pin);
}

/* computeBondForceJ2nonlinearIso for all particles in order, constitutive.c:689-863 (plmode 5); state slot [0] in place.
 * SY(x) is the macro of lpm.h:50, whose argument is not parenthesised: SY(a + dl) = 620 + 3300 (1 - exp(-0.4 a + dl)). */
void oracle_j2_iso_force_range(int ii0, int ii1, int N, int nn, double V, double J2_C, const double *Ce, const int *type, const double *xyz,
                               const int *neighbors, const int *nsign, const int *nbi, const int *nb, const double *L0, const double *Kn,
                               const double *Tv, const double *broken, const double *w, double *dLp0, double *beta0 /* [N][6] */,
                               double *alpha0, double *dlambda, double *ddLp, double *dL, double *dL_ave, double *dLt, double *TdLt, double *csx,
                               double *csy, double *csz, double *F, double *Pin)
{
    (void)N;
    int *tmp = (int *)malloc(sizeof(int) * (nn + 1));
    for (int ii = ii0; ii < ii1; ii++) {
        const int cnt = star_list(ii, nn, neighbors, broken, nb, tmp);
        for (int k = 0; k < cnt; k++)
            geom_row(tmp[k], nn, xyz, neighbors, nsign, nbi, L0, dLp0 + (long)tmp[k] * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
        for (int k = 0; k < cnt; k++) {
            const int i = tmp[k];
            const double J2_V = V * nb[i] / nn;
            double st[6] = {0, 0, 0, 0, 0, 0};
            for (int j = 0; j < nbi[i]; j++) {
                const long e = (long)i * nn + j;
                F[e] = 2.0 * Kn[e] * dL[e] + TdLt[2 * i + nsign[e]] + Tv[e] * dLt[2 * i + nsign[e]];
                F[e] *= broken[e];
                st[0] += 0.5 / J2_V * L0[e] * F[e] * csx[e] * csx[e];
                st[1] += 0.5 / J2_V * L0[e] * F[e] * csy[e] * csy[e];
                st[2] += 0.5 / J2_V * L0[e] * F[e] * csz[e] * csz[e];
                st[3] += 0.5 / J2_V * L0[e] * F[e] * csy[e] * csz[e];
                st[4] += 0.5 / J2_V * L0[e] * F[e] * csx[e] * csz[e];
                st[5] += 0.5 / J2_V * L0[e] * F[e] * csx[e] * csy[e];
            }
            const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
            for (int j = 0; j < 3; j++)
                st[j] -= temp;
            for (int j = 0; j < 6; j++)
                st[j] -= beta0[6 * i + j];
            double seq = 0.0;
            for (int j = 0; j < 6; j++)
                seq += (j < 3 ? 1.0 : 2.0) * st[j] * st[j];
            seq = sqrt(3.0 / 2.0 * seq);
            double dl = 0.0;
            double yield_func = seq - (620.0 + 3300.0 * (1.0 - exp(-0.4 * alpha0[i])));
            if (yield_func > 0.0) {
                double a = 0.0, b = 1.0, ya = yield_func;
                while ((b - a) > 1e-4) {
                    dl = (a + b) / 2.0;
                    yield_func = seq - 1.5 * dl * (2.0 * Ce[3 * type[i] + 2] + J2_C) - (620.0 + 3300.0 * (1.0 - exp(-0.4 * alpha0[i] + dl)));
                    if (yield_func * ya < 0.0) {
                        b = dl;
                    } else {
                        a = dl;
                        ya = yield_func;
                    }
                }
            }
            dlambda[i] = dl;
            alpha0[i] += dlambda[i];
            double dpl[6];
            for (int j = 0; j < 6; j++) {
                dpl[j] = dl * 1.5 * st[j] / seq;
                beta0[6 * i + j] += J2_C * dpl[j];
            }
            for (int j = 0; j < nbi[i]; j++) {
                const long e = (long)i * nn + j;
                ddLp[e] = L0[e] * (dpl[0] * csx[e] * csx[e] + dpl[1] * csy[e] * csy[e] + dpl[2] * csz[e] * csz[e] + 2 * dpl[3] * csy[e] * csz[e] +
                                   2 * dpl[4] * csx[e] * csz[e] + 2 * dpl[5] * csx[e] * csy[e]);
                dLp0[e] += ddLp[e];
            }
        }
        for (int k = 0; k < cnt; k++)
            geom_row(tmp[k], nn, xyz, neighbors, nsign, nbi, L0, dLp0 + (long)tmp[k] * nn, broken, Tv, dL, dLt, TdLt, NULL, NULL, NULL);
        const int i = ii;
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], s = nsign[e];
            for (int jj = 0; jj < nn; jj++)
                if (neighbors[(long)nj * nn + jj] == i)
                    dL_ave[e] = 0.5 * (dL[e] + dL[(long)nj * nn + jj]);
            F[e] = 2.0 * Kn[e] * dL_ave[e] + 0.5 * (TdLt[2 * i + s] + TdLt[2 * nj + s]) + 0.5 * Tv[e] * (dLt[2 * i + s] + dLt[2 * nj + s]);
            F[e] *= w[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
    }
    free(tmp);
}

void oracle_j2_iso_force(int N, int nn, double V, double J2_C, const double *Ce, const int *type, const double *xyz, const int *neighbors,
                         const int *nsign, const int *nbi, const int *nb, const double *L0, const double *Kn, const double *Tv,
                         const double *broken, const double *w, double *dLp0, double *beta0 /* [N][6] */, double *alpha0, double *dlambda,
                         double *ddLp, double *dL, double *dL_ave, double *dLt, double *TdLt, double *csx, double *csy, double *csz, double *F,
                         double *Pin)
{
    oracle_j2_iso_force_range(0, N, N, nn, V, J2_C, Ce, type, xyz, neighbors, nsign, nbi, nb, L0, Kn, Tv, broken, w, dLp0, beta0, alpha0, dlambda,
                              ddLp, dL, dL_ave, dLt, TdLt, csx, csy, csz, F, Pin);
}

/* updateDuctileDamageBwiseLocal, constitutive.c:1607-1695; pairs = newly broken (i, neighbour) in logging order */
int oracle_damage_local_bondwise(int N, int nn, double thr, double Ac, const int *neighbors, const int *nbi, const double *triax,
                                 const double *dlambda, double *dloc0, double *broken, double *dD0, double *w, int *nb, int *pairs, int max_pairs)
{
    for (int i = 0; i < N; i++) {
        const double f = (1.0 + Ac * triax[i]);
        if (f > 0.0 && dloc0[i] <= thr)
            dloc0[i] += f * dlambda[i];
        else if (dloc0[i] > 1.0)
            dloc0[i] = 1.0;
    }
    int k = 0;
    for (int i = 0; i < N; i++) {
        nb[i] = nbi[i];
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e];
            dD0[e] = 0.5 * (dloc0[i] + dloc0[nj]);
            if (dD0[e] > thr && broken[e] > EPS) {
                dD0[e] = 1.0;
                broken[e] = 0.0;
                for (int jj = 0; jj < nn; jj++)
                    if (neighbors[(long)nj * nn + jj] == i) {
                        dD0[(long)nj * nn + jj] = 1.0;
                        broken[(long)nj * nn + jj] = 0.0;
                    }
                if (k < max_pairs) {
                    pairs[2 * k] = i;
                    pairs[2 * k + 1] = nj;
                }
                k++;
            }
            if (broken[e] <= EPS)
                nb[i] -= 1;
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (fabs(broken[e]) < EPS || nb[i] == 0 || nb[neighbors[e]] == 0)
                dD0[e] = 1.0;
            w[e] = 1.0 - dD0[e];
        }
    return k;
}

/* updateDuctileDamagePwiseLocal, constitutive.c:1529-1579 (a law the reference's dispatcher keeps commented out,
 * :155): local accumulation per particle; a particle that passes the threshold is set to 1 and loses all its bonds in
 * both directions (logged as a single index); then damage_D = MAX of the two end values.  `list` receives the
 * particles, returns their number. */
int oracle_damage_local_particlewise(int N, int nn, double thr, double Ac, const int *neighbors, const int *nbi, const double *triax,
                                     const double *dlambda, double *dloc0, double *broken, double *dD0, double *w, int *list, int max_list)
{
    int k = 0;
    for (int i = 0; i < N; i++) {
        const double f = (1.0 + Ac * triax[i]);
        if (f > 0.0 && dloc0[i] <= thr)
            dloc0[i] += f * dlambda[i];
        if (dloc0[i] > thr && fabs(dloc0[i] - 1.0) > EPS) {
            dloc0[i] = 1.0;
            for (int j = 0; j < nbi[i]; j++) {
                const int nj = neighbors[(long)i * nn + j];
                broken[(long)i * nn + j] = 0.0;
                for (int jj = 0; jj < nn; jj++)
                    if (neighbors[(long)nj * nn + jj] == i)
                        broken[(long)nj * nn + jj] = 0.0;
            }
            if (k < max_list)
                list[k] = i;
            k++;
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const double a = dloc0[i], b = dloc0[neighbors[e]];
            dD0[e] = a < b ? b : a; /* MAX(x,y) ((x) < (y) ? (y) : (x)), lpm.h:46 */
            w[e] = 1.0 - dD0[e];
        }
    return k;
}

/* updateDuctileDamageBwiseNonlocal, constitutive.c:1698-1753 (also commented out in the dispatcher, :156): the
 * nonlocal average runs over the particle's own bond list with DAM_PHI(distance_initial) (lpm.h:51), the particle
 * itself enters with weight particle_volume (no phi); a bond breaks when the mean of its two end values passes the
 * threshold; every direction of a bond is visited, counted and logged on its own. */
int oracle_damage_nonlocal_bondwise(int N, int nn, double L, double thr, double Ac, double V, const int *neighbors, const int *nbi,
                                    const double *L0, const double *dlambda, const double *triax, double *Dn, double *broken, double *dD0,
                                    double *w, int *pairs, int max_pairs)
{
    for (int i = 0; i < N; i++) {
        if (Dn[i] > thr) {
            if (Dn[i] > 1.0)
                Dn[i] = 1.0;
            continue;
        }
        double DdotLocal = 0;
        double f = (1.0 + Ac * triax[i]);
        if (f > 0.0)
            DdotLocal = dlambda[i] * (1.0 + Ac * triax[i]);
        double Ddot = DdotLocal * V;
        double A = V;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e];
            DdotLocal = 0;
            f = (1.0 + Ac * triax[nj]);
            if (f > 0.0)
                DdotLocal = dlambda[nj] * (1.0 + Ac * triax[nj]);
            const double x = L0[e];
            const double phi = 1.0 / L / sqrt(2 * PI) * exp(-0.5 * x * x / L / L);
            Ddot += DdotLocal * phi * V;
            A += phi * V;
        }
        if (Ddot > 0.0)
            Dn[i] += 1.0 / A * Ddot;
    }
    int k = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (0.5 * (Dn[i] + Dn[neighbors[e]]) > thr)
                if (fabs(broken[e]) > EPS) {
                    broken[e] = 0.0;
                    dD0[e] = 1.0;
                    if (k < max_pairs) {
                        pairs[2 * k] = i;
                        pairs[2 * k + 1] = neighbors[e];
                    }
                    k++;
                }
        }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            if (fabs(broken[e]) > EPS)
                dD0[e] = 0.5 * (Dn[i] + Dn[neighbors[e]]);
            w[e] = 1.0 - dD0[e];
        }
    return k;
}

/* calcKnTv, stiffness.c:11-268: KnTve[type] = alpha * M_lattice * (C11, C12, C44)^T (alpha = radius for the 3-D lattices,
 * 1 in 2-D; the dgemm of oracle/shim: s = sum_p a[p] b[p], then alpha * s), then per bond by shell; the simple-cubic
 * lattice averages the two end particles' types (:194-200), the others use type[i] only.  lattice: 0 square, 1 hexagon,
 * 2 SC, 3 FCC, 4 BCC.  KnTve is [ntype][3] here (the hexagon uses 2 of the 3). */
void oracle_calc_kntv(int lattice, int N, int nn, int ntype, double radius, const double *Ce, const int *type, const int *neighbors,
                      const int *nsign, const int *nbi, double *KnTve, double *Kn, double *Tv)
{
    const double s2 = sqrt(2.0), s3 = sqrt(3.0);
    const double M0[9] = {1 / 2.0, -1 / 2.0, 0.0, 0.0, 0.0, 1 / 2.0, 0.0, 1.0 / 12.0, -1.0 / 12.0};
    const double M1[4] = {s3 / 12.0, -s3 / 12.0, -s3 / 144.0, s3 / 48.0};
    const double M2[9] = {1, -1, -1, 0, 0, 1, 0, 1.0 / 18.0, -1.0 / 18.0};
    const double M3[9] = {0, 0, s2, s2 / 4.0, -s2 / 4.0, -s2 / 4.0, 0, s2 / 24.0, -s2 / 24.0};
    const double M4[9] = {0., 0., s3, 1. / s3, -1. / s3, 0., 0., s3 / 14.0, s3 / 14.0};
    const double *M = lattice == 0 ? M0 : lattice == 1 ? M1 : lattice == 2 ? M2 : lattice == 3 ? M3 : M4;
    const int d = lattice == 1 ? 2 : 3;
    const double alpha = lattice >= 2 ? radius : 1.0;
    for (int k = 0; k < ntype; k++)
        for (int i = 0; i < d; i++) {
            double s = 0.0;
            for (int p = 0; p < d; p++)
                s += M[i * d + p] * Ce[3 * k + p];
            KnTve[3 * k + i] = alpha * s;
        }
    const int tv = lattice == 1 ? 1 : 2; /* column of KnTve that holds Tv */
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int sh = nsign[e], ti = type[i];
            if (lattice == 1) {
                Kn[e] = KnTve[3 * ti];
                Tv[e] = KnTve[3 * ti + 1];
            } else if (sh == 0 || sh == 1) {
                if (lattice == 2) {
                    const int tj = type[neighbors[e]];
                    Kn[e] = 0.5 * (KnTve[3 * ti + sh] + KnTve[3 * tj + sh]);
                    Tv[e] = 0.5 * (KnTve[3 * ti + tv] + KnTve[3 * tj + tv]);
                } else {
                    Kn[e] = KnTve[3 * ti + sh];
                    Tv[e] = KnTve[3 * ti + tv];
                }
            }
        }
}

/* ------------------------------------------------------------------ crystal plasticity (plmode 1)
 * computeCab, constitutive.c:1864-1917: Cab[i][a][b] = -d(RSS_a)/d(gamma_b) of the lattice stress, per particle. */
void oracle_cp_cab(int N, int nn, int S, double V, const int *nsign, const int *nb, const int *nbi, const double *Kn, const double *Tv,
                   const double *dist, const double *csx, const double *csy, const double *csz, const double *cx0, const double *cy0,
                   const double *cz0, const double *broken, const double *sch /* [S][6] */, double *Cab /* [N][S*S] */)
{
#define PROJ(e, q) (csx[e] * csx[e] * sch[6 * (q)] + csy[e] * csy[e] * sch[6 * (q) + 1] + csz[e] * csz[e] * sch[6 * (q) + 2] + \
                    csy[e] * csz[e] * sch[6 * (q) + 3] + csx[e] * csz[e] * sch[6 * (q) + 4] + csx[e] * csy[e] * sch[6 * (q) + 5])
    for (int i = 0; i < N; i++)
        for (int m = 0; m < S; m++)
            for (int n = 0; n < S; n++) {
                double LSum[2] = {0, 0};
                for (int j = 0; j < nbi[i]; j++) {
                    const long e = (long)i * nn + j;
                    LSum[nsign[e]] += Tv[e] * dist[e] * PROJ(e, n);
                }
                double acc = 0.;
                for (int j = 0; j < nbi[i]; j++) {
                    const long e = (long)i * nn + j;
                    const double dF = -2.0 * Kn[e] * dist[e] * PROJ(e, n) - 2.0 * LSum[nsign[e]];
                    const double of = opp_flag(i, j, nn, nn, nb, nbi, cx0, cy0, cz0, broken);
                    acc += -of / V * dist[e] * dF * PROJ(e, m);
                }
                Cab[(long)i * S * S + m * S + n] = acc;
            }
#undef PROJ
}

/* row-major LU with partial pivoting exactly as oracle/shim's LAPACKE_dgesv (first maximal pivot); returns info */
static int lu_solve_shim(int n, double *a, double *b)
{
    int info = 0;
    for (int k = 0; k < n; k++) {
        int p = k;
        double amax = fabs(a[k * n + k]);
        for (int i = k + 1; i < n; i++) {
            const double v = fabs(a[i * n + k]);
            if (v > amax) {
                amax = v;
                p = i;
            }
        }
        if (a[p * n + k] == 0.0) {
            if (info == 0)
                info = k + 1;
            continue;
        }
        if (p != k) {
            for (int j = 0; j < n; j++) {
                const double t = a[k * n + j];
                a[k * n + j] = a[p * n + j];
                a[p * n + j] = t;
            }
            const double t = b[k];
            b[k] = b[p];
            b[p] = t;
        }
        const double piv = a[k * n + k];
        for (int i = k + 1; i < n; i++) {
            const double l = a[i * n + k] / piv;
            a[i * n + k] = l;
            if (l != 0.0) {
                for (int j = k + 1; j < n; j++)
                    a[i * n + j] -= l * a[k * n + j];
                b[i] -= l * b[k];
            }
        }
    }
    if (info != 0)
        return info;
    for (int i = n - 1; i >= 0; i--) {
        double t = b[i];
        for (int c = i + 1; c < n; c++)
            t -= a[i * n + c] * b[c];
        b[i] = t / a[i * n + i];
    }
    return 0;
}

/* computeBondForceCPMiehe, constitutive.c:866-1396, the part that is evaluated ONCE per particle (the reference memoises
 * it in state_v[] and lets every star that contains the particle reuse it, :946-959): trial resolved shear stresses on the
 * S slip systems, active-set outer loop (<= cp_maxloop), inner Newton on D dgamma = r (<= 20 iterations, tol 1e-4) with
 * tanh/cosh hardening and power-law viscosity.  dL, dLt, TdLt, cs* hold the TRIAL geometry (slot-[0] plastic stretch).
 * Writes the increments (ddLp, cp_dA, cp_dgy, cp_dA_single), cp_Jact, cp_RSS, pl_flag and the slot-[2] state.
 * Returns 0, or i + 1 if the slip Jacobian of particle i is singular (the reference exits there, :1216-1221). */
/* the return map of the particles rows[0..nrows) (rows == NULL: all N particles) */
static int cp_return_map_rows(int nrows, const int *rows, int N, int nn, int S, double V, double h0, double taus, double tau0, double q,
                              double eta, double pp, double maxloop, double dtime, const int *nsign, const int *nb, const int *nbi,
                              const double *Kn, const double *Tv, const double *w, const double *broken, const double *L0, const double *cx0,
                              const double *cy0, const double *cz0, const double *dL, const double *dLt, const double *TdLt, const double *csx,
                              const double *csy, const double *csz, const double *sch, const double *Cab, const double *dLp0,
                              const double *gy0 /* [N][S] */, const double *A0 /* [N] */, const double *As0 /* [N][S] */, double *ddLp,
                              double *dA, double *dgy, double *dAs, int *Jact, double *RSS, int *pl_flag, double *dLp2, double *gy2, double *A2,
                              double *As2)
{
    double *gam = (double *)malloc(sizeof(double) * S), *r = (double *)malloc(sizeof(double) * S), *rhs = (double *)malloc(sizeof(double) * S);
    double *D = (double *)malloc(sizeof(double) * S * S), *xgy = (double *)malloc(sizeof(double) * S), *yf = (double *)malloc(sizeof(double) * S);
    double *xdL = (double *)malloc(sizeof(double) * nn);
    int bad = 0;
    const int count = rows ? nrows : N;
    for (int qq = 0; qq < count && !bad; qq++) {
        const int i = rows ? rows[qq] : qq;
        const long b0 = (long)i * nn;
        double st[6] = {0}, xt[2] = {dLt[2 * i], dLt[2 * i + 1]}, xT[2] = {TdLt[2 * i], TdLt[2 * i + 1]};
        for (int j = 0; j < nbi[i]; j++)
            xdL[j] = dL[b0 + j];
#define STRESS()                                                                                                      \
    do {                                                                                                              \
        memset(st, 0, sizeof st);                                                                                     \
        for (int j = 0; j < nbi[i]; j++) {                                                                            \
            const long e = b0 + j;                                                                                    \
            double Fij = 2.0 * Kn[e] * xdL[j] + xT[nsign[e]] + Tv[e] * xt[nsign[e]];                                   \
            Fij *= w[e];                                                                                              \
            const double of = opp_flag(i, j, nn, nn, nb, nbi, cx0, cy0, cz0, broken);                                  \
            st[0] += of / V * L0[e] * Fij * csx[e] * csx[e];                                                          \
            st[1] += of / V * L0[e] * Fij * csy[e] * csy[e];                                                          \
            st[2] += of / V * L0[e] * Fij * csz[e] * csz[e];                                                          \
            st[3] += of / V * L0[e] * Fij * csy[e] * csz[e];                                                          \
            st[4] += of / V * L0[e] * Fij * csx[e] * csz[e];                                                          \
            st[5] += of / V * L0[e] * Fij * csx[e] * csy[e];                                                          \
        }                                                                                                             \
    } while (0)
#define RSS_OF(m) (st[0] * sch[6 * (m)] + st[1] * sch[6 * (m) + 1] + st[2] * sch[6 * (m) + 2] + st[3] * sch[6 * (m) + 3] + \
                   st[4] * sch[6 * (m) + 4] + st[5] * sch[6 * (m) + 5])
        STRESS();
        double tmax = 0.0;
        for (int m = 0; m < S; m++) {
            RSS[(long)i * S + m] = RSS_OF(m);
            yf[m] = RSS[(long)i * S + m] - gy0[(long)i * S + m];
            if (yf[m] > tmax)
                tmax = yf[m];
            xgy[m] = gy0[(long)i * S + m];
        }
        double xA = A0[i];
        for (int m = 0; m < S; m++)
            gam[m] = 0.0;
        if (tmax <= EPS) {
            for (int j = 0; j < nbi[i]; j++)
                ddLp[b0 + j] = 0.0;
            dA[i] = 0.0;
            for (int m = 0; m < S; m++) {
                Jact[(long)i * S + m] = 0;
                dgy[(long)i * S + m] = 0.0;
                dAs[(long)i * S + m] = 0.0;
            }
        } else {
            int *J = Jact + (long)i * S;
            int outer = 0;
            pl_flag[i] = 1;
            memset(J, 0, sizeof(int) * S);
            for (;;) {
                outer++;
                double norm_r = 1.0;
                memset(gam, 0, sizeof(double) * S);
                memset(r, 0, sizeof(double) * S);
                memset(rhs, 0, sizeof(double) * S);
                memset(D, 0, sizeof(double) * S * S);
                int inner = 0;
                do {
                    inner++;
                    double dpl[6] = {0};
                    for (int s2 = 0; s2 < S; s2++)
                        for (int c = 0; c < 6; c++)
                            dpl[c] += J[s2] * gam[s2] * sch[6 * s2 + c];
                    xt[0] = xt[1] = xT[0] = xT[1] = 0;
                    for (int j = 0; j < nbi[i]; j++) {
                        const long e = b0 + j;
                        ddLp[e] = L0[e] * (dpl[0] * csx[e] * csx[e] + dpl[1] * csy[e] * csy[e] + dpl[2] * csz[e] * csz[e] + dpl[3] * csy[e] * csz[e] +
                                           dpl[4] * csx[e] * csz[e] + dpl[5] * csx[e] * csy[e]);
                        ddLp[e] *= broken[e];
                        xdL[j] = dL[e] - ddLp[e];
                        xt[nsign[e]] += xdL[j];
                        xT[nsign[e]] += Tv[e] * xdL[j];
                    }
                    STRESS();
                    dA[i] = 0.0;
                    for (int s2 = 0; s2 < S; s2++)
                        dA[i] += gam[s2];
                    xA = A0[i] + dA[i];
                    const double h_hat = h0 / pow(cosh(h0 * xA / (taus - tau0)), 2.0);
                    const double h_hatp = -2.0 * h0 * h0 / (taus - tau0) * tanh(h0 * xA / (taus - tau0)) * h_hat;
                    for (int a = 0; a < S; a++) {
                        double t1 = 0.0;
                        for (int b = 0; b < S; b++) {
                            const double hab = a == b ? h_hat : q * h_hat;
                            t1 += J[b] * hab * gam[b];
                        }
                        dgy[(long)i * S + a] = J[a] * t1;
                        xgy[a] = gy0[(long)i * S + a] + dgy[(long)i * S + a];
                    }
                    for (int m = 0; m < S; m++) {
                        const double t1 = pow(1. + gam[m] * eta / dtime, 1. / pp);
                        RSS[(long)i * S + m] = RSS_OF(m);
                        r[m] = J[m] * (RSS[(long)i * S + m] - xgy[m] * t1);
                        rhs[m] = r[m];
                    }
                    for (int m = 0; m < S; m++)
                        for (int n = 0; n < S; n++) {
                            if (J[m] == 1 && J[n] == 1) {
                                double h_star = 0.0;
                                for (int d = 0; d < S; d++) {
                                    double hd = 0.0;
                                    if (m == d && n == d)
                                        hd = h_hat + h_hatp * gam[d];
                                    else if (m == d && n != d)
                                        hd = h_hatp * gam[d];
                                    else if (m != d && n == d)
                                        hd = q * (h_hat + h_hatp * gam[d]);
                                    else
                                        hd = q * h_hatp * gam[d];
                                    h_star += J[d] * hd;
                                }
                                const double t1 = xgy[m] * (eta / pp / dtime * pow(1. + eta * gam[m] / dtime, (1. - pp) / pp));
                                const double t2 = h_star * pow(1. + eta * gam[m] / dtime, (1. / pp));
                                D[m * S + n] = m == n ? Cab[(long)i * S * S + m * S + n] + t1 + t2 : Cab[(long)i * S * S + m * S + n] + t2;
                            } else if (m == n)
                                D[m * S + n] = 1.0;
                        }
                    if (lu_solve_shim(S, D, rhs) > 0) {
                        bad = i + 1;
                        break;
                    }
                    for (int m = 0; m < S; m++)
                        gam[m] += J[m] * rhs[m];
                    double nr2 = 0.0;
                    for (int m = 0; m < S; m++)
                        nr2 += r[m] * r[m];
                    norm_r = sqrt(nr2);
                } while (norm_r > 1e-4 && inner < 20);
                if (bad)
                    break;
                int minI = -1, maxI = -1;
                double minY = 0.0, maxY = 0.0;
                for (int m = 0; m < S; m++) {
                    yf[m] = RSS[(long)i * S + m] - xgy[m];
                    if (J[m] == 1 && gam[m] <= 0.0 && yf[m] < minY) {
                        minY = yf[m];
                        minI = m;
                    }
                }
                if (minI != -1) {
                    J[minI] = 0;
                    continue;
                }
                for (int m = 0; m < S; m++)
                    if (J[m] == 0 && yf[m] > 0.0 && yf[m] > maxY) {
                        maxY = yf[m];
                        maxI = m;
                    }
                if (maxI != -1) {
                    J[maxI] = 1;
                    if (outer < maxloop)
                        continue;
                }
                break;
            }
            for (int m = 0; m < S; m++)
                dAs[(long)i * S + m] = J[m] * gam[m];
        }
        /* slot [2]: what the star owner stores for itself (:1371-1379) */
        for (int j = 0; j < nn; j++) {
            double xd = dLp0[b0 + j];
            if (j < nbi[i])
                xd += ddLp[b0 + j];
            dLp2[b0 + j] = broken[b0 + j] * xd;
        }
        for (int m = 0; m < S; m++) {
            gy2[(long)i * S + m] = xgy[m];
            As2[(long)i * S + m] = As0[(long)i * S + m] + dAs[(long)i * S + m];
        }
        A2[i] = xA;
#undef STRESS
#undef RSS_OF
    }
    free(gam); free(r); free(rhs); free(D); free(xgy); free(yf); free(xdL);
    return bad;
}

int oracle_cp_return_map(int N, int nn, int S, double V, double h0, double taus, double tau0, double q, double eta, double pp, double maxloop,
                         double dtime, const int *nsign, const int *nb, const int *nbi, const double *Kn, const double *Tv, const double *w,
                         const double *broken, const double *L0, const double *cx0, const double *cy0, const double *cz0, const double *dL,
                         const double *dLt, const double *TdLt, const double *csx, const double *csy, const double *csz, const double *sch,
                         const double *Cab, const double *dLp0, const double *gy0 /* [N][S] */, const double *A0 /* [N] */,
                         const double *As0 /* [N][S] */, double *ddLp, double *dA, double *dgy, double *dAs, int *Jact, double *RSS,
                         int *pl_flag, double *dLp2, double *gy2, double *A2, double *As2)
{
    return cp_return_map_rows(0, NULL, N, nn, S, V, h0, taus, tau0, q, eta, pp, maxloop, dtime, nsign, nb, nbi, Kn, Tv, w, broken, L0, cx0, cy0, cz0,
                              dL, dLt, TdLt, csx, csy, csz, sch, Cab, dLp0, gy0, A0, As0, ddLp, dA, dgy, dAs, Jact, RSS, pl_flag, dLp2, gy2, A2, As2);
}

/* computeBondForceCPMiehe(ii) called on its own, constitutive.c:866-1396, with the memo state_v (:946-959): the geometry of
 * ii's star with the slot-[0] plastic stretch (:912-935); star members still flagged REUSE the increments an earlier call
 * left, the others are return-mapped and flagged (:938-1317); the star's geometry with dLp[0] + ddLp (:1320-1341); the force
 * pass of ii over the arrays as they are (:1343-1361); slot [2] of ii (:1371-1379).  Returns the reference's exit condition
 * (singular slip Jacobian) as a non-zero value. */
int oracle_cp_particle(int ii, int N, int nn, int S, double V, double h0, double taus, double tau0, double q, double eta, double pp,
                       double maxloop, double dtime, const double *xyz, const int *neighbors, const int *nsign, const int *nb, const int *nbi,
                       const double *Kn, const double *Tv, const double *w, const double *broken, const double *L0, const double *cx0,
                       const double *cy0, const double *cz0, const double *sch, const double *Cab, const double *dLp0, const double *gy0,
                       const double *A0, const double *As0, int *state_v, double *dL, double *dLt, double *TdLt, double *csx, double *csy,
                       double *csz, double *ddLp, double *dA, double *dgy, double *dAs, int *Jact, double *RSS, int *pl_flag, double *dL_ave,
                       double *F, double *Pin, double *dLp2, double *gy2, double *A2, double *As2)
{
    int *star = (int *)malloc(sizeof(int) * (nn + 1)), *rows0 = (int *)malloc(sizeof(int) * (nn + 1));
    const int cnt = star_list(ii, nn, neighbors, broken, nb, star);
    for (int k = 0; k < cnt; k++)
        geom_row(star[k], nn, xyz, neighbors, nsign, nbi, L0, dLp0 + (long)star[k] * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
    int n0 = 0;
    for (int k = 0; k < cnt; k++)
        if (state_v[star[k]] == 0) {
            state_v[star[k]] = 1;
            rows0[n0++] = star[k];
        }
    const int ii_fresh = n0 > 0 && rows0[0] == ii;
    /* slot [2] is only stored for ii (:1371-1379): the other star members' go to scratch */
    double *t_dLp2 = (double *)malloc(sizeof(double) * (size_t)N * nn), *t_gy2 = (double *)malloc(sizeof(double) * (size_t)N * S);
    double *t_A2 = (double *)malloc(sizeof(double) * N), *t_As2 = (double *)malloc(sizeof(double) * (size_t)N * S);
    int bad = 0;
    if (n0 > 0)
        bad = cp_return_map_rows(n0, rows0, N, nn, S, V, h0, taus, tau0, q, eta, pp, maxloop, dtime, nsign, nb, nbi, Kn, Tv, w, broken, L0, cx0,
                                 cy0, cz0, dL, dLt, TdLt, csx, csy, csz, sch, Cab, dLp0, gy0, A0, As0, ddLp, dA, dgy, dAs, Jact, RSS, pl_flag,
                                 t_dLp2, t_gy2, t_A2, t_As2);
    double *xd = (double *)malloc(sizeof(double) * nn);
    for (int k = 0; k < cnt; k++) {
        const int i = star[k];
        for (int j = 0; j < nn; j++)
            xd[j] = dLp0[(long)i * nn + j] + (j < nbi[i] ? ddLp[(long)i * nn + j] : 0.0);
        geom_row(i, nn, xyz, neighbors, nsign, nbi, L0, xd, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
        if (i == ii && !ii_fresh)
            for (int j = 0; j < nn; j++)
                dLp2[(long)i * nn + j] = broken[(long)i * nn + j] * xd[j];
    }
    {
        const int i = ii;
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], s = nsign[e];
            for (int jj = 0; jj < nn; jj++)
                if (neighbors[(long)nj * nn + jj] == i)
                    dL_ave[e] = 0.5 * (dL[e] + dL[(long)nj * nn + jj]);
            F[e] = 2.0 * Kn[e] * dL_ave[e] + 0.5 * (TdLt[2 * i + s] + TdLt[2 * nj + s]) + 0.5 * Tv[e] * (dLt[2 * i + s] + dLt[2 * nj + s]);
            F[e] *= w[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
        if (ii_fresh) {
            memcpy(dLp2 + (long)i * nn, t_dLp2 + (long)i * nn, sizeof(double) * nn);
            memcpy(gy2 + (long)i * S, t_gy2 + (long)i * S, sizeof(double) * S);
            memcpy(As2 + (long)i * S, t_As2 + (long)i * S, sizeof(double) * S);
            A2[i] = t_A2[i];
        } else {
            for (int m = 0; m < S; m++) {
                gy2[(long)i * S + m] = gy0[(long)i * S + m] + dgy[(long)i * S + m];
                As2[(long)i * S + m] = As0[(long)i * S + m] + dAs[(long)i * S + m];
            }
            A2[i] = A0[i] + dA[i];
        }
    }
    free(star); free(rows0); free(t_dLp2); free(t_gy2); free(t_A2); free(t_As2); free(xd);
    return bad;
}

/* updateBrittleDamage, constitutive.c:1437-1526 (plmode 6): bonds with dL / L0 >= critical_bstrain are candidates (the
 * reference holds at most 400, :1444); all of them break if there are <= nbreak, else the nbreak largest after the
 * reference's (non-stable) shell sort.  Returns the CANDIDATE count like the reference; `pairs` = broken (i, neighbour). */
int oracle_damage_brittle(int N, int nn, double crit, int nbreak, const int *neighbors, const int *nbi, const double *dL, const double *L0,
                          double *broken, double *dD0, double *w, int *pairs, int max_pairs)
{
    enum { CAP = 400 };
    int bi[CAP], bj[CAP];
    double bs[CAP];
    int k = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < nbi[i]; j++) {
            const double ave = dL[(long)i * nn + j] / L0[(long)i * nn + j];
            if (ave >= crit) {
                if (k >= CAP)
                    return -1; /* the reference overruns b_cr[] here */
                bi[k] = i;
                bj[k] = j;
                bs[k] = ave;
                k++;
            }
        }
    int first = 0;
    if (k > nbreak) {
        for (int r = k / 2; r >= 1; r = r / 2)
            for (int i = r; i < k; ++i) {
                const int ti = bi[i], tj = bj[i];
                const double tb = bs[i];
                int j = i - r;
                while (j >= 0 && bs[j] > tb) {
                    bs[j + r] = bs[j];
                    bi[j + r] = bi[j];
                    bj[j + r] = bj[j];
                    j = j - r;
                }
                bs[j + r] = tb;
                bi[j + r] = ti;
                bj[j + r] = tj;
            }
        first = k - nbreak;
    }
    for (int i = first; i < k; i++) {
        const long e = (long)bi[i] * nn + bj[i];
        dD0[e] = 1.0;
        w[e] = 0.0;
        broken[e] = 0.0;
        if (i - first < max_pairs) {
            pairs[2 * (i - first)] = bi[i];
            pairs[2 * (i - first) + 1] = neighbors[e];
        }
    }
    return k;
}

/* ------------------------------------------------------------------ the per-particle law entry points, called
 * outside the dispatcher: computeBondForceElastic(ii) constitutive.c:228-283 (law 6), computeBondForceIncrementalUpdating(ii)
 * :167-225 (law 4), computeBondForceJ2mixedLinear3D(ii) :466-686 (law 0).  Literal: temporaries per member of the star
 * (ii + neighbours across intact bonds), geometry / return map for every member, force pass and state for ii only. */
void oracle_particle_law(int law, int ii, int N, int nn, double V, double J2_H, double J2_xi, const double *Ce, const int *type,
                         const double *sigmay, const double *xyz, const double *xyz_temp, const int *neighbors, const int *nsign, const int *nbi,
                         const int *nb, const double *L0, const double *cx0, const double *cy0, const double *cz0, const double *Kn,
                         const double *Tv, const double *broken, const double *w, const double *F_temp, const double *dLp0, const double *beta0,
                         const double *alpha0, double *dL, double *dLt, double *TdLt, double *csx, double *csy, double *csz, double *ddL,
                         double *ddLt, double *TddLt, double *ddLp, int *pl_flag, double *dL_ave, double *F, double *Pin, double *stress,
                         double *dlambda, double *dLp2, double *beta2, double *alpha2)
{
    int *star = (int *)malloc(sizeof(int) * (nn + 1));
    const int ns = star_list(ii, nn, neighbors, broken, nb, star);
    if (law == 4) {
        for (int k = 0; k < ns; k++) {
            const int i = star[k];
            ddLt[2 * i] = ddLt[2 * i + 1] = TddLt[2 * i] = TddLt[2 * i + 1] = 0;
            for (int j = 0; j < nbi[i]; j++) {
                const long e = (long)i * nn + j;
                const int nj = neighbors[e];
                const double ax = xyz_temp[3 * i] - xyz_temp[3 * nj], ay = xyz_temp[3 * i + 1] - xyz_temp[3 * nj + 1],
                             az = xyz_temp[3 * i + 2] - xyz_temp[3 * nj + 2];
                const double dis0 = sqrt(ax * ax + ay * ay + az * az);
                const double bx = xyz[3 * i] - xyz[3 * nj], by = xyz[3 * i + 1] - xyz[3 * nj + 1], bz = xyz[3 * i + 2] - xyz[3 * nj + 2];
                const double dis1 = sqrt(bx * bx + by * by + bz * bz);
                ddL[e] = broken[e] * (dis1 - dis0);
                ddLt[2 * i + nsign[e]] += ddL[e];
                TddLt[2 * i + nsign[e]] += Tv[e] * ddL[e];
            }
        }
        const int i = ii;
        Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int nj = neighbors[e], sg = nsign[e];
            F[e] = F_temp[e] + 2.0 * Kn[e] * ddL[e] + 0.5 * (TddLt[2 * i + sg] + TddLt[2 * nj + sg]) + 0.5 * Tv[e] * (ddLt[2 * i + sg] + ddLt[2 * nj + sg]);
            F[e] *= broken[e];
            Pin[3 * i] += csx[e] * F[e];
            Pin[3 * i + 1] += csy[e] * F[e];
            Pin[3 * i + 2] += csz[e] * F[e];
        }
        free(star);
        return;
    }
    /* plastic stretch the geometry passes subtract: slot [0] of every member, advanced by the return map (law 0) */
    double *xdLp = (double *)malloc(sizeof(double) * ns * nn), *xbeta = (double *)malloc(sizeof(double) * ns * 6);
    double *xalpha = (double *)malloc(sizeof(double) * ns), *xdl = (double *)calloc(ns, sizeof(double));
    for (int k = 0; k < ns; k++) {
        const int i = star[k];
        for (int j = 0; j < nn; j++)
            xdLp[k * nn + j] = dLp0[(long)i * nn + j];
        for (int q = 0; q < 6; q++)
            xbeta[k * 6 + q] = beta0[6 * i + q];
        xalpha[k] = alpha0[i];
    }
    for (int k = 0; k < ns; k++)
        geom_row(star[k], nn, xyz, neighbors, nsign, nbi, L0, xdLp + k * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
    if (law == 0) {
        for (int k = 0; k < ns; k++) {
            const int i = star[k];
            double st[6] = {0}, dpl[6] = {0};
            for (int j = 0; j < nbi[i]; j++) {
                const long e = (long)i * nn + j;
                double Fij = 2.0 * Kn[e] * dL[e] + TdLt[2 * i + nsign[e]] + Tv[e] * dLt[2 * i + nsign[e]];
                Fij *= w[e];
                const double of = opp_flag(i, j, nn, nn, nb, nbi, cx0, cy0, cz0, broken);
                st[0] += of / V * L0[e] * Fij * csx[e] * csx[e];
                st[1] += of / V * L0[e] * Fij * csy[e] * csy[e];
                st[2] += of / V * L0[e] * Fij * csz[e] * csz[e];
                st[3] += of / V * L0[e] * Fij * csy[e] * csz[e];
                st[4] += of / V * L0[e] * Fij * csx[e] * csz[e];
                st[5] += of / V * L0[e] * Fij * csx[e] * csy[e];
            }
            const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
            for (int q = 0; q < 3; q++)
                st[q] -= temp;
            for (int q = 0; q < 6; q++)
                st[q] -= xbeta[k * 6 + q];
            double seq = 0.0;
            for (int q = 0; q < 6; q++)
                seq += (q < 3 ? 1.0 : 2.0) * st[q] * st[q];
            seq = sqrt(3.0 / 2.0 * seq);
            const double yf = seq - (sigmay[i] + (1.0 - J2_xi) * J2_H * xalpha[k]);
            if (yf > 0.0) {
                pl_flag[i] = 1;
                xdl[k] = yf / (3 * Ce[3 * type[i] + 2] + J2_H);
            }
            xalpha[k] += xdl[k];
            for (int q = 0; q < 6; q++)
                if (fabs(seq) > EPS) {
                    dpl[q] = xdl[k] * 1.5 * st[q] / seq;
                    xbeta[k * 6 + q] += 2. / 3. * J2_xi * J2_H * dpl[q];
                }
            for (int j = 0; j < nbi[i]; j++) {
                const long e = (long)i * nn + j;
                ddLp[e] = L0[e] * (dpl[0] * csx[e] * csx[e] + dpl[1] * csy[e] * csy[e] + dpl[2] * csz[e] * csz[e] + 2 * dpl[3] * csy[e] * csz[e] +
                                   2 * dpl[4] * csx[e] * csz[e] + 2 * dpl[5] * csx[e] * csy[e]);
                ddLp[e] *= broken[e];
                xdLp[k * nn + j] += ddLp[e];
            }
        }
        for (int k = 0; k < ns; k++)
            geom_row(star[k], nn, xyz, neighbors, nsign, nbi, L0, xdLp + k * nn, broken, Tv, dL, dLt, TdLt, csx, csy, csz);
    }
    const int i = ii;
    if (law == 0)
        for (int q = 0; q < 6; q++)
            stress[6 * i + q] = 0.0;
    Pin[3 * i] = Pin[3 * i + 1] = Pin[3 * i + 2] = 0.0;
    for (int j = 0; j < nbi[i]; j++) {
        const long e = (long)i * nn + j;
        const int nj = neighbors[e], sg = nsign[e];
        double stretch = dL[e];
        if (law == 0) {
            for (int jj = 0; jj < nn; jj++)
                if (neighbors[(long)nj * nn + jj] == i)
                    dL_ave[e] = 0.5 * (dL[e] + dL[(long)nj * nn + jj]);
            stretch = dL_ave[e];
        }
        F[e] = 2.0 * Kn[e] * stretch + 0.5 * (TdLt[2 * i + sg] + TdLt[2 * nj + sg]) + 0.5 * Tv[e] * (dLt[2 * i + sg] + dLt[2 * nj + sg]);
        F[e] *= law == 0 ? w[e] : broken[e];
        Pin[3 * i] += csx[e] * F[e];
        Pin[3 * i + 1] += csy[e] * F[e];
        Pin[3 * i + 2] += csz[e] * F[e];
    }
    if (law == 0) {
        for (int j = 0; j < nn; j++)
            dLp2[(long)i * nn + j] = broken[(long)i * nn + j] * xdLp[j];
        for (int q = 0; q < 6; q++)
            beta2[6 * i + q] = xbeta[q];
        alpha2[i] = xalpha[0];
        dlambda[i] = xdl[0];
    }
    free(star);
    free(xdLp);
    free(xbeta);
    free(xalpha);
    free(xdl);
}

/* computeStrain, lpm_basic.c:127-249, with the LU of oracle/shim's LAPACKE_dgesv (row-major, first maximal pivot;
 * right-hand side untouched when singular, as LAPACK's dgesv) */
void oracle_compute_strain(int N, int nn, int dim, const double *xyz0, const int *neighbors, const int *nsign, const int *nbi,
                           const double *L0, const double *dL, double *strain /* [N][6] */)
{
    const int ns = 3 * (dim - 1);
    for (int i = 0; i < N; i++) {
        double r4[3][3][3][3] = {{{{0}}}}, r2[3][3] = {{0}};
        for (int j = 0; j < nbi[i]; j++) {
            const long e = (long)i * nn + j;
            const int s = nsign[e], nj = neighbors[e];
            if (s != 0 && s != 1)
                continue;
            const double wgt = s == 0 ? 0.1 : 0.9;
            double dx[3];
            for (int k = 0; k < dim; k++)
                dx[k] = xyz0[3 * nj + k] - xyz0[3 * i + k];
            for (int k = 0; k < dim; k++)
                for (int n = 0; n < dim; n++) {
                    for (int m = 0; m < dim; m++)
                        for (int l = 0; l < dim; l++)
                            r4[k][n][m][l] += wgt * dx[k] / L0[e] * dx[n] / L0[e] * dx[m] / L0[e] * dx[l] / L0[e];
                    r2[k][n] += wgt * dL[e] / L0[e] * dx[k] / L0[e] * dx[n] / L0[e];
                }
        }
        double a[36], b[6], b_in[6];
        int ii = 0;
        for (int j = 0; j < dim; j++)
            for (int k = 0; k < dim; k++)
                for (int m = 0; m < dim; m++)
                    for (int l = 0; l < dim; l++)
                        if (m <= l && j <= k)
                            a[ii++] = r4[m][l][j][k];
        ii = 0;
        for (int j = 0; j < dim; j++)
            for (int k = 0; k < dim; k++)
                if (j <= k) {
                    b[ii] = b_in[ii] = r2[j][k];
                    ii++;
                }
        int info = 0;
        for (int k = 0; k < ns; k++) {
            int p = k;
            double amax = fabs(a[k * ns + k]);
            for (int r = k + 1; r < ns; r++)
                if (fabs(a[r * ns + k]) > amax) {
                    amax = fabs(a[r * ns + k]);
                    p = r;
                }
            if (a[p * ns + k] == 0.0) {
                if (info == 0)
                    info = k + 1;
                continue;
            }
            if (p != k) {
                for (int c = 0; c < ns; c++) {
                    const double t = a[k * ns + c];
                    a[k * ns + c] = a[p * ns + c];
                    a[p * ns + c] = t;
                }
                const double t = b[k];
                b[k] = b[p];
                b[p] = t;
            }
            for (int r = k + 1; r < ns; r++) {
                const double l = a[r * ns + k] / a[k * ns + k];
                a[r * ns + k] = l;
                if (l != 0.0) {
                    for (int c = k + 1; c < ns; c++)
                        a[r * ns + c] -= l * a[k * ns + c];
                    b[r] -= l * b[k];
                }
            }
        }
        if (info != 0) {
            if (i == 0)
                continue;
            memcpy(b, b_in, sizeof(double) * ns);
        } else {
            for (int r = ns - 1; r >= 0; r--) {
                double t = b[r];
                for (int c = r + 1; c < ns; c++)
                    t -= a[r * ns + c] * b[c];
                b[r] = t / a[r * ns + r];
            }
        }
        if (dim == 2) {
            strain[6 * i] = b[0];
            strain[6 * i + 5] = b[1];
            strain[6 * i + 1] = b[2];
        } else {
            strain[6 * i] = b[0];
            strain[6 * i + 5] = b[1];
            strain[6 * i + 4] = b[2];
            strain[6 * i + 1] = b[3];
            strain[6 * i + 3] = b[4];
            strain[6 * i + 2] = b[5];
        }
    }
}
