// lpmb_cp.cu -- rate-dependent crystal plasticity (Miehe 2001) return map and the elastic RSS-sensitivity
// matrix Cab (compiled -fmad=false).
//
// Replaces, in the reference:
//   computeBondForceCPMiehe(ii)   src/constitutive.c:866-1396   (plmode 1; run serially there, :114-117, with the
//                                 per-call memo state_v so that every particle's update is computed once)
//   computeCab()                  src/constitutive.c:1864-1917
//
// One thread per particle runs the whole active-set iteration of its particle: trial resolved shear stresses,
// outer loop over the active set (<= cp_maxloop), inner Newton on the slips gamma (<= MAXSMALL=20 iterations,
// tolerance TOLITER) with the dense Jacobian solved by an in-thread row-major LU with partial pivoting (the pivot
// rule of LAPACKE_dgesv: first row of maximal |a|), tanh/cosh hardening, power-law viscosity.  The loop is
// branch-divergent by nature (each particle has its own active set); the Jacobian lives in local memory.  The
// increments (ddLp, cp_dgy, cp_dA, cp_dA_single) and the slot-[2] state are written exactly as the reference
// leaves them, including its one-iteration lag between gamma and ddLp at loop exit.  pow/cosh/tanh differ from
// glibc by <= 1-2 ulp, so this law is compared at 1e-9, not bit for bit; Cab is + - * / only and is bit-exact.
#include "lpmb_internal.cuh"

#define CPT 64
#define MAXSMALL 20  // include/lpm.h:38

struct CPParams {
    double V, h0, taus, tau0, q, eta, p, dtime, maxloop;
    int S;
};

__device__ __forceinline__ double cp_opp_flag(int nb_i, int nn, int o, const double *__restrict__ broken, size_t Np, int i)
{
    if (nb_i == nn)
        return 0.5;
    if (o < 0)
        return 1.0;
    return broken[(size_t)o * Np + i] <= LPMB_EPS ? 1.0 : 0.5;
}

// row-major LU with partial pivoting, one right-hand side; returns k+1 on an exactly zero pivot (LAPACK info)
template <int SMAX>
__device__ int dgesv_rowmajor(int n, double *a, double *b)
{
    for (int k = 0; k < n; k++) {
        int p = k;
        double amax = fabs(a[k * SMAX + k]);
        for (int i = k + 1; i < n; i++) {
            const double v = fabs(a[i * SMAX + k]);
            if (v > amax) {
                amax = v;
                p = i;
            }
        }
        if (a[p * SMAX + k] == 0.0)
            return k + 1;
        if (p != k) {
            for (int j = 0; j < n; j++) {
                const double t = a[k * SMAX + j];
                a[k * SMAX + j] = a[p * SMAX + j];
                a[p * SMAX + j] = t;
            }
            const double t = b[k];
            b[k] = b[p];
            b[p] = t;
        }
        const double piv = a[k * SMAX + k];
        for (int i = k + 1; i < n; i++) {
            const double l = a[i * SMAX + k] / piv;
            a[i * SMAX + k] = l;
            if (l != 0.0) {
                for (int j = k + 1; j < n; j++)
                    a[i * SMAX + j] -= l * a[k * SMAX + j];
                b[i] -= l * b[k];
            }
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < n; c++)
            s -= a[i * SMAX + c] * b[c];
        b[i] = s / a[i * SMAX + i];
    }
    return 0;
}

template <int SMAX>
__global__ void __launch_bounds__(CPT)
cp_miehe_kernel(int N, int Np, int nn, CPParams P, const int *__restrict__ nbi_g, const int *__restrict__ nb_g, const signed char *__restrict__ nsign,
                const signed char *__restrict__ opp, const double *__restrict__ schmid /* [S][6] */, const double *__restrict__ Kn,
                const double *__restrict__ Tv, const double *__restrict__ w, const double *__restrict__ broken, const double *__restrict__ L0,
                const double *__restrict__ dL, const double *__restrict__ dLt, const double *__restrict__ TdLt, const double *__restrict__ csx,
                const double *__restrict__ csy, const double *__restrict__ csz, const double *__restrict__ dLp0, const double *__restrict__ gy0,
                const double *__restrict__ A0, const double *__restrict__ As0, const double *__restrict__ Cab, double *__restrict__ dLp2,
                double *__restrict__ gy2, double *__restrict__ A2, double *__restrict__ As2, double *__restrict__ ddLp, double *__restrict__ RSS,
                int *__restrict__ Jact, double *__restrict__ dgy, double *__restrict__ dA, double *__restrict__ dAs, int *__restrict__ pl_flag,
                int *__restrict__ err)
{
    const int i = blockIdx.x * CPT + threadIdx.x;
    if (i >= N)
        return;
    const size_t Npz = Np;
    const int S = P.S;
    const int n = nbi_g[i], nb_i = nb_g[i];
    double gamma[SMAX], r[SMAX], rrhs[SMAX], xgy[SMAX], yf[SMAX], D[SMAX * SMAX];
    signed char jact[SMAX];
    double st[6] = {0, 0, 0, 0, 0, 0};

    // trial stress from the trial elastic stretches (constitutive.c:961-995)
    {
        const double t0 = dLt[i], t1 = dLt[Npz + i], T0 = TdLt[i], T1 = TdLt[Npz + i];
        for (int j = 0; j < n; j++) {
            const size_t e = (size_t)j * Npz + i;
            const int s = nsign[e];
            double Fij = 2.0 * Kn[e] * dL[e] + (s ? T1 : T0) + Tv[e] * (s ? t1 : t0);
            Fij *= w[e];
            const double of = cp_opp_flag(nb_i, nn, opp[e], broken, Npz, i);
            const double cx = csx[e], cy = csy[e], cz = csz[e];
            const double pre = of / P.V * L0[e] * Fij;
            st[0] += pre * cx * cx;
            st[1] += pre * cy * cy;
            st[2] += pre * cz * cz;
            st[3] += pre * cy * cz;
            st[4] += pre * cx * cz;
            st[5] += pre * cx * cy;
        }
    }
    double temp_max = 0.0;
    for (int m = 0; m < S; m++) {
        const double *sm = schmid + 6 * m;
        const double rss = st[0] * sm[0] + st[1] * sm[1] + st[2] * sm[2] + st[3] * sm[3] + st[4] * sm[4] + st[5] * sm[5];
        RSS[(size_t)m * Npz + i] = rss;
        xgy[m] = gy0[(size_t)m * Npz + i];
        yf[m] = rss - xgy[m];
        if (yf[m] > temp_max)
            temp_max = yf[m];
        gamma[m] = 0.0;
        jact[m] = 0;
    }
    double xA = A0[i];
    double dA_i = 0.0;

    if (temp_max <= LPMB_EPS) {  // elastic step (constitutive.c:1011-1022)
        for (int j = 0; j < n; j++)
            ddLp[(size_t)j * Npz + i] = 0.0;
        for (int m = 0; m < S; m++)
            dgy[(size_t)m * Npz + i] = 0.0;
    } else {
        int niter_outer = 0;
        pl_flag[i] = 1;
        for (;;) {  // label_outer
            niter_outer++;
            double norm_r = 1.0;
            for (int m = 0; m < S; m++) {
                gamma[m] = 0.0;
                r[m] = 0.0;
                rrhs[m] = 0.0;
            }
            int niter_inner = 0;
            do {
                niter_inner++;
                double dpl[6] = {0, 0, 0, 0, 0, 0};
                for (int s = 0; s < S; s++) {
                    const double *sm = schmid + 6 * s;
                    const double jg = jact[s] * gamma[s];
#pragma unroll
                    for (int q = 0; q < 6; q++)
                        dpl[q] += jg * sm[q];
                }
                // updated elastic stretches (engineering shear: no factor 2, constitutive.c:1062-1075)
                double t[2] = {0, 0}, T[2] = {0, 0};
                for (int j = 0; j < n; j++) {
                    const size_t e = (size_t)j * Npz + i;
                    const double cx = csx[e], cy = csy[e], cz = csz[e];
                    double dd = L0[e] * (dpl[0] * cx * cx + dpl[1] * cy * cy + dpl[2] * cz * cz + dpl[3] * cy * cz + dpl[4] * cx * cz + dpl[5] * cx * cy);
                    dd *= broken[e];
                    ddLp[e] = dd;
                    const double xd = dL[e] - dd;
                    const int s = nsign[e];
                    t[s] += xd;
                    T[s] += Tv[e] * xd;
                }
                for (int q = 0; q < 6; q++)
                    st[q] = 0.0;
                for (int j = 0; j < n; j++) {
                    const size_t e = (size_t)j * Npz + i;
                    const int s = nsign[e];
                    const double xd = dL[e] - ddLp[e];
                    double Fij = 2.0 * Kn[e] * xd + T[s] + Tv[e] * t[s];
                    Fij *= w[e];
                    const double of = cp_opp_flag(nb_i, nn, opp[e], broken, Npz, i);
                    const double cx = csx[e], cy = csy[e], cz = csz[e];
                    const double pre = of / P.V * L0[e] * Fij;
                    st[0] += pre * cx * cx;
                    st[1] += pre * cy * cy;
                    st[2] += pre * cz * cz;
                    st[3] += pre * cy * cz;
                    st[4] += pre * cx * cz;
                    st[5] += pre * cx * cy;
                }
                dA_i = 0.0;
                for (int s = 0; s < S; s++)
                    dA_i += gamma[s];
                xA = A0[i] + dA_i;
                const double ch = cosh(P.h0 * xA / (P.taus - P.tau0));
                const double h_hat = P.h0 / (ch * ch);
                const double h_hatp = -2.0 * P.h0 * P.h0 / (P.taus - P.tau0) * tanh(P.h0 * xA / (P.taus - P.tau0)) * h_hat;
                for (int a = 0; a < S; a++) {
                    double term1 = 0.0;
                    for (int b = 0; b < S; b++) {
                        const double hab = (a == b) ? h_hat : P.q * h_hat;
                        term1 += jact[b] * hab * gamma[b];
                    }
                    const double dg = jact[a] * term1;
                    dgy[(size_t)a * Npz + i] = dg;
                    xgy[a] = gy0[(size_t)a * Npz + i] + dg;
                }
                for (int m = 0; m < S; m++) {
                    const double *sm = schmid + 6 * m;
                    const double term1 = pow(1. + gamma[m] * P.eta / P.dtime, 1. / P.p);
                    const double rss = st[0] * sm[0] + st[1] * sm[1] + st[2] * sm[2] + st[3] * sm[3] + st[4] * sm[4] + st[5] * sm[5];
                    RSS[(size_t)m * Npz + i] = rss;
                    yf[m] = rss;  // keep the RSS of this iteration for the yield functions below
                    r[m] = jact[m] * (rss - xgy[m] * term1);
                    rrhs[m] = r[m];
                }
                for (int m = 0; m < S; m++) {
                    // the two powers of the viscous term depend on the row only: once per active row instead of once per
                    // active PAIR (same arguments -> the same bits; pow is ~half of this kernel's instructions otherwise)
                    double pw_a = 0.0, pw_b = 0.0;
                    if (jact[m] == 1) {
                        pw_a = pow(1. + P.eta * gamma[m] / P.dtime, (1. - P.p) / P.p);
                        pw_b = pow(1. + P.eta * gamma[m] / P.dtime, (1. / P.p));
                    }
                    for (int nn2 = 0; nn2 < S; nn2++) {
                        double v = (m == nn2) ? 1.0 : 0.0;
                        if (jact[m] == 1 && jact[nn2] == 1) {
                            double h_star = 0.0;
                            for (int d = 0; d < S; d++) {
                                double hd;
                                if (m == d && nn2 == d)
                                    hd = h_hat + h_hatp * gamma[d];
                                else if (m == d && nn2 != d)
                                    hd = h_hatp * gamma[d];
                                else if (m != d && nn2 == d)
                                    hd = P.q * (h_hat + h_hatp * gamma[d]);
                                else
                                    hd = P.q * h_hatp * gamma[d];
                                h_star += jact[d] * hd;
                            }
                            const double term1 = xgy[m] * (P.eta / P.p / P.dtime * pw_a);
                            const double term2 = h_star * pw_b;
                            const double cab = Cab[(size_t)(m * S + nn2) * Npz + i];
                            v = (m == nn2) ? cab + term1 + term2 : cab + term2;
                        }
                        D[m * SMAX + nn2] = v;
                    }
                }
                if (dgesv_rowmajor<SMAX>(S, D, rrhs) != 0) {
                    atomicExch(err, i + 1);  // the reference prints and exit(1)s (constitutive.c:1216-1221)
                    return;
                }
                for (int m = 0; m < S; m++)
                    gamma[m] += jact[m] * rrhs[m];
                double s2 = 0.0;
                for (int m = 0; m < S; m++)
                    s2 += r[m] * r[m];
                norm_r = sqrt(s2);
            } while (norm_r > LPMB_TOLITER && niter_inner < MAXSMALL);

            // active-set update (constitutive.c:1250-1304)
            int minIndex = -1, maxIndex = -1;
            double minYield = 0.0, maxYield = 0.0;
            for (int m = 0; m < S; m++) {
                yf[m] = yf[m] - xgy[m];  // cp_RSS - xcp_gy
                if (jact[m] == 1 && gamma[m] <= 0.0 && yf[m] < minYield) {
                    minYield = yf[m];
                    minIndex = m;
                }
            }
            if (minIndex != -1) {
                jact[minIndex] = 0;
                continue;  // goto label_outer
            }
            for (int m = 0; m < S; m++)
                if (jact[m] == 0 && yf[m] > 0.0 && yf[m] > maxYield) {
                    maxYield = yf[m];
                    maxIndex = m;
                }
            if (maxIndex != -1) {
                jact[maxIndex] = 1;
                if (niter_outer < P.maxloop)
                    continue;  // goto label_outer
            }
            break;  // label_outside
        }
    }
    // label_outside (constitutive.c:1307-1317) + slot-[2] state (1371-1379)
    for (int j = 0; j < nn; j++) {
        const size_t e = (size_t)j * Npz + i;
        double xd = dLp0[e];
        if (j < n)
            xd += ddLp[e];
        dLp2[e] = broken[e] * xd;
    }
    for (int s = 0; s < S; s++) {
        const size_t e = (size_t)s * Npz + i;
        const double das = jact[s] * gamma[s];
        dAs[e] = das;
        As2[e] = As0[e] + das;
        gy2[e] = xgy[s];
        Jact[e] = jact[s];
    }
    dA[i] = dA_i;
    A2[i] = xA;
}

// ---------------------------------------------------------------------------------------------
// Warp-per-particle version of the same return map (S <= 32 slip systems, nn <= 32; param `cp_warp`, default 1).
//
// The thread-per-particle kernel above leaves a B200 nearly empty at BASELINE config 4's size (6 912 particles = 216
// warps on 148 SMs), keeps its 24 x 24 Jacobian in local memory and makes the 32 particles of a warp wait for each
// other's active-set histories.  Here ONE WARP owns a particle:
//   * lane j owns bond j (stretch update, bond force, the six stress products), lane m owns slip system m (resolved
//     shear stress, slip, hardening, residual, the viscous powers -- the pow calls of all systems side by side);
//   * everything is restricted to the ACTIVE SET (warp-uniform bit mask): an inactive system has gamma = +0, residual 0
//     and an identity row / column in the Jacobian, so it contributes exact zeros to every sum of the serial algorithm
//     and drops out of the LU untouched (its diagonal 1 is the only non-zero of its column: never a pivot candidate for
//     another column, multiplier 0 for every row).  The Jacobian is therefore assembled and factorised as the compact
//     A x A system of the active rows in shared memory (A = 1..8 in practice), with LAPACK's pivot rule (first row of
//     maximal |a|), the eliminations of one column side by side, and the back substitution in the serial order;
//   * sums the reference accumulates in a fixed order (shell sums and stress over the bonds; dpl, dA, hardening, h_star,
//     residual norm over the slip systems) are accumulated in that order by one lane per sum from shared memory.
// Every stored value sees the same operations on the same operands as in the kernel above, except for additions of an
// exact zero that are skipped: the two kernels agree bit for bit up to the sign of a zero
// (tests/test_cp_gpu.py::test_warp_kernel_equals_thread_kernel), and both sit within 1e-9 of the reference (glibc's pow /
// cosh / tanh differ from CUDA's by <= 1-2 ulp).  Non-finite intermediate values (the reference produces NaNs in one BCC
// iteration and rolls them back) are not reproduced operation by operation here; 48-system lattices use the kernel above.
// ---------------------------------------------------------------------------------------------
#define CPW_WARPS 4
#define CPW_S 24  // slip systems handled per warp (compact Jacobian CPW_S x CPW_S)
__device__ __forceinline__ double cpw_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

struct CpwSmem {
    double D[CPW_S][CPW_S + 1];   // compact Jacobian of the active rows
    double rhs[CPW_S], sol[CPW_S];
    double PA[32], PB[32], PC[32], PE[32];   // h_star terms by slip system
    double xgy[32], pwa[32], pwb[32];
    double xd[32], tvx[32], pr[6][32];       // per-bond values of the current iteration
    int sg[32];
    int act[32];
};

__global__ void __launch_bounds__(32 * CPW_WARPS)
cp_miehe_warp_kernel(int N, int Np, int nn, CPParams P, const int *__restrict__ nbi_g, const int *__restrict__ nb_g,
                     const signed char *__restrict__ nsign, const signed char *__restrict__ opp, const double *__restrict__ schmid /* [S][6] */,
                     const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ w, const double *__restrict__ broken,
                     const double *__restrict__ L0, const double *__restrict__ dL, const double *__restrict__ dLt, const double *__restrict__ TdLt,
                     const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz, const double *__restrict__ dLp0,
                     const double *__restrict__ gy0, const double *__restrict__ A0, const double *__restrict__ As0, const double *__restrict__ Cab,
                     double *__restrict__ dLp2, double *__restrict__ gy2, double *__restrict__ A2, double *__restrict__ As2, double *__restrict__ ddLp,
                     double *__restrict__ RSS, int *__restrict__ Jact, double *__restrict__ dgy, double *__restrict__ dA, double *__restrict__ dAs,
                     int *__restrict__ pl_flag, int *__restrict__ err)
{
    __shared__ CpwSmem smem_all[CPW_WARPS];
    const int lane = threadIdx.x & 31;
    CpwSmem &W = smem_all[threadIdx.x >> 5];
    const int i = blockIdx.x * CPW_WARPS + (threadIdx.x >> 5);
    if (i >= N)
        return;  // the whole warp leaves; only warp-level synchronisation below
    const size_t Npz = Np;
    const int S = P.S;
    const int n = nbi_g[i], nb_i = nb_g[i];
    const bool ml = lane < S, bl = lane < n;
    const int m = lane;
    // bond `lane` of the particle
    double cx = 0, cy = 0, cz = 0, bL0 = 0, bbrk = 0, bdL = 0, bKn = 0, bTv = 0, bw = 0, bof = 0;
    int bsg = 0;
    if (bl) {
        const size_t e = (size_t)lane * Npz + i;
        cx = csx[e];
        cy = csy[e];
        cz = csz[e];
        bL0 = L0[e];
        bbrk = broken[e];
        bdL = dL[e];
        bKn = Kn[e];
        bTv = Tv[e];
        bw = w[e];
        bof = cp_opp_flag(nb_i, nn, opp[e], broken, Npz, i);
        bsg = nsign[e];
    }
    W.sg[lane] = bsg;
    double sm[6] = {0, 0, 0, 0, 0, 0};
    if (ml) {
#pragma unroll
        for (int q = 0; q < 6; q++)
            sm[q] = schmid[6 * m + q];
    }
    // stress from per-bond stretches xd and shell sums t / T: lanes 0..5 add the six products over the bonds in order
    double st[6];
    auto stress_from = [&](double xd, double t0, double t1, double T0, double T1) {
        double Fij = 2.0 * bKn * xd + (bsg ? T1 : T0) + bTv * (bsg ? t1 : t0);
        Fij *= bw;
        const double pre = bof / P.V * bL0 * Fij;
        W.pr[0][lane] = pre * cx * cx;
        W.pr[1][lane] = pre * cy * cy;
        W.pr[2][lane] = pre * cz * cz;
        W.pr[3][lane] = pre * cy * cz;
        W.pr[4][lane] = pre * cx * cz;
        W.pr[5][lane] = pre * cx * cy;
        __syncwarp();
        double acc = 0.0;
        if (lane < 6)
            for (int j = 0; j < n; j++)
                acc += W.pr[lane][j];
#pragma unroll
        for (int q = 0; q < 6; q++)
            st[q] = cpw_shfl(acc, q);
        __syncwarp();
    };
    // trial stress from the trial elastic stretches (constitutive.c:961-995)
    stress_from(bdL, dLt[i], dLt[Npz + i], TdLt[i], TdLt[Npz + i]);
    double rss = st[0] * sm[0] + st[1] * sm[1] + st[2] * sm[2] + st[3] * sm[3] + st[4] * sm[4] + st[5] * sm[5];
    const double gy0m = ml ? gy0[(size_t)m * Npz + i] : 0.0;
    double xgy = gy0m;
    double yf = rss - xgy;
    double gamma = 0.0, dg_last = 0.0, mydd = 0.0;
    int jact = 0;
    // temp_max = max(0, max_m yf[m]) (a NaN never wins, as in the serial `if (yf[m] > temp_max)`)
    double temp_max = (ml && yf > 0.0) ? yf : 0.0;
    for (int o = 16; o > 0; o >>= 1)
        temp_max = fmax(temp_max, cpw_shfl(temp_max, lane ^ o));
    const double A0i = A0[i];
    double xA = A0i, dA_i = 0.0;

    if (temp_max > LPMB_EPS) {
        int niter_outer = 0;
        if (lane == 0)
            pl_flag[i] = 1;
        for (;;) {  // label_outer
            niter_outer++;
            double norm_r = 1.0;
            gamma = 0.0;
            double r = 0.0;
            // the active set of this pass: ascending list, count, my position in it
            const unsigned amask = __ballot_sync(0xffffffffu, ml && jact == 1);
            const int A = __popc(amask);
            const int mypos = __popc(amask & ((1u << lane) - 1u));
            if (jact == 1)
                W.act[mypos] = m;
            __syncwarp();
            int niter_inner = 0;
            do {
                niter_inner++;
                double dpl[6] = {0, 0, 0, 0, 0, 0};
                for (int ka = 0; ka < A; ka++) {
                    const int s = W.act[ka];
                    const double jgs = cpw_shfl(gamma, s);   // jact[s] * gamma[s] with jact[s] = 1
                    const double *ss = schmid + 6 * s;
#pragma unroll
                    for (int q = 0; q < 6; q++)
                        dpl[q] += jgs * ss[q];
                }
                // updated elastic stretches (engineering shear: no factor 2, constitutive.c:1062-1075)
                double dd = bL0 * (dpl[0] * cx * cx + dpl[1] * cy * cy + dpl[2] * cz * cz + dpl[3] * cy * cz + dpl[4] * cx * cz + dpl[5] * cx * cy);
                dd *= bbrk;
                mydd = dd;
                const double xd = bdL - dd;
                W.xd[lane] = xd;
                W.tvx[lane] = bTv * xd;
                __syncwarp();
                double acc = 0.0;
                if (lane < 4) {  // lane 0: t[0], 1: t[1], 2: T[0], 3: T[1]
                    const int shell = lane & 1;
                    for (int j = 0; j < n; j++)
                        if (W.sg[j] == shell)
                            acc += lane < 2 ? W.xd[j] : W.tvx[j];
                }
                const double t0 = cpw_shfl(acc, 0), t1 = cpw_shfl(acc, 1), T0 = cpw_shfl(acc, 2), T1 = cpw_shfl(acc, 3);
                __syncwarp();
                stress_from(xd, t0, t1, T0, T1);
                dA_i = 0.0;
                for (int ka = 0; ka < A; ka++)
                    dA_i += cpw_shfl(gamma, W.act[ka]);
                xA = A0i + dA_i;
                const double ch = cosh(P.h0 * xA / (P.taus - P.tau0));
                const double h_hat = P.h0 / (ch * ch);
                const double h_hatp = -2.0 * P.h0 * P.h0 / (P.taus - P.tau0) * tanh(P.h0 * xA / (P.taus - P.tau0)) * h_hat;
                {
                    double term1 = 0.0;
                    for (int ka = 0; ka < A; ka++) {
                        const int b = W.act[ka];
                        const double gb = cpw_shfl(gamma, b);
                        const double hab = (m == b) ? h_hat : P.q * h_hat;
                        term1 += hab * gb;   // jact[b] * hab * gamma[b] with jact[b] = 1
                    }
                    dg_last = jact * term1;
                    xgy = gy0m + dg_last;
                }
                rss = st[0] * sm[0] + st[1] * sm[1] + st[2] * sm[2] + st[3] * sm[3] + st[4] * sm[4] + st[5] * sm[5];
                yf = rss;  // keep the RSS of this iteration for the yield functions below
                double pw_a = 0.0, pw_b = 0.0;
                r = 0.0;
                if (jact == 1) {
                    const double term1 = pow(1. + gamma * P.eta / P.dtime, 1. / P.p);
                    r = rss - xgy * term1;
                    pw_a = pow(1. + P.eta * gamma / P.dtime, (1. - P.p) / P.p);
                    pw_b = pow(1. + P.eta * gamma / P.dtime, (1. / P.p));
                    const double ga = h_hat + h_hatp * gamma;
                    W.PA[m] = ga;
                    W.PB[m] = h_hatp * gamma;
                    W.PC[m] = P.q * ga;
                    W.PE[m] = P.q * h_hatp * gamma;
                    W.xgy[m] = xgy;
                    W.pwa[m] = pw_a;
                    W.pwb[m] = pw_b;
                    W.rhs[mypos] = r;
                }
                __syncwarp();
                // compact Jacobian of the active rows / columns (constitutive.c:1142-1210)
                for (int e = lane; e < A * A; e += 32) {
                    const int ka = e / A, kb = e - ka * A;
                    const int ma = W.act[ka], n2 = W.act[kb];
                    double h_star = 0.0;
                    for (int kd = 0; kd < A; kd++) {
                        const int d = W.act[kd];
                        double hd;
                        if (ma == d && n2 == d)
                            hd = W.PA[d];
                        else if (ma == d && n2 != d)
                            hd = W.PB[d];
                        else if (ma != d && n2 == d)
                            hd = W.PC[d];
                        else
                            hd = W.PE[d];
                        h_star += hd;
                    }
                    const double term1 = W.xgy[ma] * (P.eta / P.p / P.dtime * W.pwa[ma]);
                    const double term2 = h_star * W.pwb[ma];
                    const double cab = Cab[(size_t)(ma * S + n2) * Npz + i];
                    W.D[ka][kb] = (ma == n2) ? cab + term1 + term2 : cab + term2;
                }
                __syncwarp();
                // ---- LU with partial pivoting (LAPACK's pivot rule) on the compact system; lane = row
                bool singular = false;
                for (int k = 0; k < A; k++) {
                    double v = -1.0;
                    if (lane >= k && lane < A) {
                        v = fabs(W.D[lane][k]);
                        if (v != v)
                            v = lane == k ? __longlong_as_double(0x7ff0000000000000ll) : -1.0;
                    }
                    int pi = lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = cpw_shfl(v, lane ^ o);
                        const int oi = __shfl_sync(0xffffffffu, pi, lane ^ o);
                        if (ov > v || (ov == v && oi < pi)) {
                            v = ov;
                            pi = oi;
                        }
                    }
                    const int p = pi;  // warp-uniform
                    if (W.D[p][k] == 0.0) {
                        singular = true;
                        break;
                    }
                    if (p != k) {
                        if (lane < A) {
                            const double tk = W.D[k][lane];
                            W.D[k][lane] = W.D[p][lane];
                            W.D[p][lane] = tk;
                        }
                        if (lane == 0) {
                            const double tb = W.rhs[k];
                            W.rhs[k] = W.rhs[p];
                            W.rhs[p] = tb;
                        }
                        __syncwarp();
                    }
                    const double piv = W.D[k][k];
                    double l = 0.0;
                    if (lane > k && lane < A) {
                        l = W.D[lane][k] / piv;
                        W.D[lane][k] = l;
                        if (l != 0.0) {
                            for (int j = k + 1; j < A; j++)
                                W.D[lane][j] -= l * W.D[k][j];
                            W.rhs[lane] -= l * W.rhs[k];
                        }
                    }
                    __syncwarp();
                }
                if (singular) {
                    if (lane == 0)
                        atomicExch(err, i + 1);  // the reference prints and exit(1)s (constitutive.c:1216-1221)
                    return;
                }
                // back substitution in the serial order (every lane runs the same chain on broadcast reads)
                for (int ii = A - 1; ii >= 0; ii--) {
                    double sacc = W.rhs[ii];
                    for (int c2 = ii + 1; c2 < A; c2++)
                        sacc -= W.D[ii][c2] * W.sol[c2];
                    const double xi = sacc / W.D[ii][ii];
                    __syncwarp();
                    if (lane == 0)
                        W.sol[ii] = xi;
                    __syncwarp();
                }
                if (jact == 1)
                    gamma += W.sol[mypos];   // gamma[m] += jact[m] * rrhs[m]
                double s2 = 0.0;
                for (int ka = 0; ka < A; ka++) {
                    const double rs = cpw_shfl(r, W.act[ka]);
                    s2 += rs * rs;
                }
                norm_r = sqrt(s2);
                __syncwarp();
            } while (norm_r > LPMB_TOLITER && niter_inner < MAXSMALL);

            // active-set update (constitutive.c:1250-1304): first index of the most negative / most positive yield function
            yf = yf - xgy;  // cp_RSS - xcp_gy
            {
                const bool cand = ml && jact == 1 && gamma <= 0.0 && yf < 0.0;
                double v = cand ? yf : 0.0;
                int idx = cand ? lane : 64;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = cpw_shfl(v, lane ^ o);
                    const int oi = __shfl_sync(0xffffffffu, idx, lane ^ o);
                    if (oi < 64 && (idx == 64 || ov < v || (ov == v && oi < idx))) {
                        v = ov;
                        idx = oi;
                    }
                }
                if (idx < 64) {
                    if (lane == idx)
                        jact = 0;
                    continue;  // goto label_outer
                }
            }
            {
                const bool cand = ml && jact == 0 && yf > 0.0;
                double v = cand ? yf : 0.0;
                int idx = cand ? lane : 64;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = cpw_shfl(v, lane ^ o);
                    const int oi = __shfl_sync(0xffffffffu, idx, lane ^ o);
                    if (oi < 64 && (idx == 64 || ov > v || (ov == v && oi < idx))) {
                        v = ov;
                        idx = oi;
                    }
                }
                if (idx < 64) {
                    if (lane == idx)
                        jact = 1;
                    if (niter_outer < P.maxloop)
                        continue;  // goto label_outer
                }
            }
            break;  // label_outside
        }
    }
    // label_outside (constitutive.c:1307-1317) + slot-[2] state (1371-1379)
    if (lane < nn) {
        const size_t e = (size_t)lane * Npz + i;
        double xdp = dLp0[e];
        if (bl) {
            ddLp[e] = mydd;
            xdp += mydd;
        }
        dLp2[e] = broken[e] * xdp;
    }
    if (ml) {
        const size_t e = (size_t)m * Npz + i;
        RSS[e] = rss;
        dgy[e] = dg_last;
        const double das = jact * gamma;
        dAs[e] = das;
        As2[e] = As0[e] + das;
        gy2[e] = xgy;
        Jact[e] = jact;
    }
    if (lane == 0) {
        dA[i] = dA_i;
        A2[i] = xA;
    }
}

// computeCab   constitutive.c:1864-1917
__global__ void __launch_bounds__(CPT)
compute_cab_kernel(int N, int Np, int nn, int S, double V, const int *__restrict__ nbi_g, const int *__restrict__ nb_g,
                   const signed char *__restrict__ nsign, const signed char *__restrict__ opp, const double *__restrict__ schmid,
                   const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ broken, const double *__restrict__ distance,
                   const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ Cab)
{
    const int i = blockIdx.x * CPT + threadIdx.x;
    if (i >= N)
        return;
    const size_t Npz = Np;
    const int n = nbi_g[i], nb_i = nb_g[i];
    for (int m = 0; m < S; m++) {
        const double *pm = schmid + 6 * m;
        for (int n2 = 0; n2 < S; n2++) {
            const double *pn = schmid + 6 * n2;
            double LSum[2] = {0, 0};
            for (int j = 0; j < n; j++) {
                const size_t e = (size_t)j * Npz + i;
                const double cx = csx[e], cy = csy[e], cz = csz[e];
                LSum[nsign[e]] += Tv[e] * distance[e] * (cx * cx * pn[0] + cy * cy * pn[1] + cz * cz * pn[2] + cy * cz * pn[3] + cx * cz * pn[4] + cx * cy * pn[5]);
            }
            double cab = 0.;
            for (int j = 0; j < n; j++) {
                const size_t e = (size_t)j * Npz + i;
                const double cx = csx[e], cy = csy[e], cz = csz[e];
                const double dF = -2.0 * Kn[e] * distance[e] * (cx * cx * pn[0] + cy * cy * pn[1] + cz * cz * pn[2] + cy * cz * pn[3] + cx * cz * pn[4] + cx * cy * pn[5]) -
                                  2.0 * LSum[nsign[e]];
                const double of = cp_opp_flag(nb_i, nn, opp[e], broken, Npz, i);
                cab += -of / V * distance[e] * dF * (cx * cx * pm[0] + cy * cy * pm[1] + cz * cz * pm[2] + cy * cz * pm[3] + cx * cz * pm[4] + cx * cy * pm[5]);
            }
            Cab[(size_t)(m * S + n2) * Npz + i] = cab;
        }
    }
}

static int cp_need(lpmb_ctx *c, int *S_out)
{
    LPMB_REQUIRE(c->params.count("nslipSys"), LPMB_ERR_STATE, "parameter nslipSys not set");
    const int S = (int)param(c, "nslipSys");
    LPMB_REQUIRE(S > 0 && S <= 48, LPMB_ERR_UNSUPPORTED, "nslipSys=%d (1..48 supported)", S);
    LPMB_REQUIRE(c->fields.count("schmid_tensor"), LPMB_ERR_STATE, "schmid_tensor not uploaded (lpmb_set_schmid_tensor)");
    *S_out = S;
    return LPMB_OK;
}

extern "C" int lpmb_set_schmid_tensor(lpmb_ctx *c, const double *schmid, int nslipSys)
{
    LPMB_REQUIRE(c && schmid && nslipSys > 0 && nslipSys <= 48, LPMB_ERR_ARG, "lpmb_set_schmid_tensor: bad argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    c->params["nslipSys"] = nslipSys;
    auto it = c->fields.find("schmid_tensor");
    if (it != c->fields.end()) {
        cudaFree(it->second.d);
        c->fields.erase(it);
    }
    Field f;
    f.kind = FK_RAW;
    f.type = FT_F64;
    f.comps = 6;
    f.count = (size_t)nslipSys * 6;
    LPMB_CUDA(cudaMalloc(&f.d, f.count * 8));
    LPMB_H2D(c, f.d, schmid, f.count * 8);
    c->fields["schmid_tensor"] = f;
    return LPMB_OK;
}

extern "C" int lpmb_compute_cab(lpmb_ctx *c)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    int S = 0;
    LPMB_TRY(cp_need(c, &S));
    LPMB_REQUIRE(c->params.count("particle_volume"), LPMB_ERR_STATE, "particle_volume not set");
    compute_cab_kernel<<<lpmb_blocks(c->N, CPT), CPT, 0, c->stream>>>(
        c->N, c->Np, c->nn, S, param(c, "particle_volume"), fptr<int>(c, "nb_initial"), fptr<int>(c, "nb"), fptr<signed char>(c, "nsign"),
        fptr<signed char>(c, "oppslot"), (const double *)c->fields["schmid_tensor"].d, fptr<double>(c, "Kn"), fptr<double>(c, "Tv"),
        fptr<double>(c, "damage_broken"), fptr<double>(c, "distance"), fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"),
        fptr<double>(c, "cp_Cab"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

// the return-map part of computeBondForceGeneral(1, t); geometry / force / stress passes are shared with the J2 law
int lpmb_cp_return_map(lpmb_ctx *c)
{
    int S = 0;
    LPMB_TRY(cp_need(c, &S));
    CPIO io;
    io.dL = fptr<double>(c, "dL");
    io.dLt = fptr<double>(c, "dL_total");
    io.TdLt = fptr<double>(c, "TdL_total");
    io.csx = fptr<double>(c, "csx");
    io.csy = fptr<double>(c, "csy");
    io.csz = fptr<double>(c, "csz");
    io.dLp2 = fptr<double>(c, "dLp2");
    io.gy2 = fptr<double>(c, "cp_gy2");
    io.A2 = fptr<double>(c, "cp_A2");
    io.As2 = fptr<double>(c, "cp_A_single2");
    io.ddLp = fptr<double>(c, "ddLp");
    io.RSS = fptr<double>(c, "cp_RSS");
    io.Jact = fptr<int>(c, "cp_Jact");
    io.dgy = fptr<double>(c, "cp_dgy");
    io.dA = fptr<double>(c, "cp_dA");
    io.dAs = fptr<double>(c, "cp_dA_single");
    io.pl_flag = fptr<int>(c, "pl_flag");
    return lpmb_cp_return_map_io(c, &io, nullptr);
}

int lpmb_cp_return_map_io(lpmb_ctx *c, const CPIO *io, int *err_particle)
{
    int S = 0;
    LPMB_TRY(cp_need(c, &S));
    for (const char *p : {"cp_h0", "cp_taus0", "cp_tau00", "cp_q", "cp_eta", "cp_p", "cp_maxloop", "dtime", "particle_volume"})
        LPMB_REQUIRE(c->params.count(p), LPMB_ERR_STATE, "parameter %s not set", p);
    CPParams P;
    P.V = param(c, "particle_volume");
    P.h0 = param(c, "cp_h0");
    P.taus = param(c, "cp_taus0");
    P.tau0 = param(c, "cp_tau00");
    P.q = param(c, "cp_q");
    P.eta = param(c, "cp_eta");
    P.p = param(c, "cp_p");
    P.dtime = param(c, "dtime");
    P.maxloop = param(c, "cp_maxloop");
    P.S = S;
    int *d_err = reinterpret_cast<int *>(c->cg.scal ? c->cg.scal + 15 : nullptr);
    if (!d_err) {
        LPMB_TRY(lpmb_cg_alloc(c));
        d_err = reinterpret_cast<int *>(c->cg.scal + 15);
    }
    LPMB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
#define CP_ARGS                                                                                                                                   \
    c->N, c->Np, c->nn, P, fptr<int>(c, "nb_initial"), fptr<int>(c, "nb"), fptr<signed char>(c, "nsign"), fptr<signed char>(c, "oppslot"),       \
        (const double *)c->fields["schmid_tensor"].d, fptr<double>(c, "Kn"), fptr<double>(c, "Tv"), fptr<double>(c, "damage_w"),                  \
        fptr<double>(c, "damage_broken"), fptr<double>(c, "distance_initial"), io->dL, io->dLt, io->TdLt, io->csx, io->csy, io->csz,              \
        fptr<double>(c, "dLp0"), fptr<double>(c, "cp_gy0"), fptr<double>(c, "cp_A0"), fptr<double>(c, "cp_A_single0"), fptr<double>(c, "cp_Cab"), \
        io->dLp2, io->gy2, io->A2, io->As2, io->ddLp, io->RSS, io->Jact, io->dgy, io->dA, io->dAs, io->pl_flag, d_err
    if (S <= 24 && c->nn <= 32 && param(c, "cp_warp", 1.0) != 0.0)
        cp_miehe_warp_kernel<<<lpmb_blocks(c->N, CPW_WARPS), 32 * CPW_WARPS, 0, c->stream>>>(CP_ARGS);
    else if (S <= 24)
        cp_miehe_kernel<24><<<lpmb_blocks(c->N, CPT), CPT, 0, c->stream>>>(CP_ARGS);
    else
        cp_miehe_kernel<48><<<lpmb_blocks(c->N, CPT), CPT, 0, c->stream>>>(CP_ARGS);
#undef CP_ARGS
    LPMB_LAUNCH_CHECK(c);
    int err = 0;
    LPMB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    if (err_particle) {
        *err_particle = err;
        return LPMB_OK;
    }
    LPMB_REQUIRE(err == 0, LPMB_ERR_STATE, "crystal plasticity: singular slip Jacobian at particle %d (the reference exits here, constitutive.c:1216-1221)",
                 err - 1);
    return LPMB_OK;
}
