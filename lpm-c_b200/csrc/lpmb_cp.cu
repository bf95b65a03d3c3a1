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
    if (S <= 24)
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
