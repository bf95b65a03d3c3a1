// lpmb_bond.cu -- bond-wise constitutive update (compiled with -fmad=false: strict IEEE, so every
// + - * / sqrt reproduces the reference's gcc -ffp-contract=off evaluation bit for bit).
//
// Replaces, in the reference:
//   computeBondForceGeneral(plmode,t)        src/constitutive.c:88-146   (dispatch + computeStress + switchStateV(2))
//   computeBondForceElastic(ii)              src/constitutive.c:228-283  (plmode 6)
//   computeBondForceIncrementalUpdating(ii)  src/constitutive.c:167-225  (plmode 4, predictor)
//   computeBondForceJ2mixedLinear3D(ii)      src/constitutive.c:466-686  (plmode 0)
//   computeBondForceJ2energyReturnMap(ii,t)  src/constitutive.c:286-463  (plmode 3)
//   computeBondForceJ2nonlinearIso(ii)       src/constitutive.c:689-863  (plmode 5)
//   computeStress()                          src/lpm_basic.c:53-125
//   computedL()                              src/lpm_basic.c:252-291
//   computeStrain()                          src/lpm_basic.c:127-249
//   switchStateV(flag)                       src/constitutive.c:10-85
//   updateRR()                               src/stiffness.c:519-534
//   updateCrack()                            src/constitutive.c:1399-1434
//
// Structure.  The reference evaluates, for every particle ii, the geometry / return map of ii AND
// of each of its neighbours (19x redundant); each of those evaluations is a pure function of the
// slot-[0] state and xyz, so here every particle is evaluated once, in passes:
//   geometry -> [return map -> geometry] -> force (+Pin) -> stress -> state switch
// one thread per particle, per-bond arrays slot-major ([slot][Np]) so that a warp reads 32
// consecutive doubles per slot; neighbour quantities (positions, dilatation sums, mirror-bond
// stretch) are gathered through L2.  No atomics: every particle owns its outputs.
#include "lpmb_internal.cuh"

#define BT 128  // threads per block for the particle-parallel kernels

struct BondView {
    int N, Np, nn, dim;
    const int *nbr;             // [nn][Np]
    const signed char *nsign;   // [nn][Np]
    const signed char *mirror;  // [nn][Np] slot of the reverse bond in the neighbour's list
    const signed char *opp;     // [nn][Np] slot of the geometrically opposite bond (or -1)
    const int *nbi;             // nb_initial
    int *nb;
    const double *xyz;          // [3][Np]
};

__global__ void fill_int_kernel(int *__restrict__ p, int n, int v);

static int make_view(lpmb_ctx *c, BondView &v)
{
    v.N = c->N;
    v.Np = c->Np;
    v.nn = c->nn;
    v.dim = c->dim;
    v.nbr = fptr<int>(c, "neighbors");
    v.nsign = fptr<signed char>(c, "nsign");
    v.mirror = fptr<signed char>(c, "mirror");
    v.opp = fptr<signed char>(c, "oppslot");
    v.nbi = fptr<int>(c, "nb_initial");
    v.nb = fptr<int>(c, "nb");
    v.xyz = fptr<double>(c, "xyz");
    LPMB_REQUIRE(v.nbr && v.nsign && v.mirror && v.opp && v.nbi && v.nb && v.xyz, LPMB_ERR_STATE, "topology fields missing");
    return LPMB_OK;
}

// ---------------------------------------------------------------------------------------------
// topology derived data: nb_initial, mirror slot, opposite-bond slot
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
derive_topology_kernel(int N, int Np, int nn, const int *__restrict__ nbr, int *__restrict__ nbi, int *__restrict__ nb,
                       signed char *__restrict__ mirror)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= N)
        return;
    int n = 0;
    for (int j = 0; j < nn; j++) {
        const int nj = nbr[(size_t)j * Np + i];
        signed char mj = -1;
        if (nj >= 0) {
            n++;
            // constitutive.c:654-658: the slot jj with neighbors[neighbors[i][j]][jj] == i (last match wins)
            for (int jj = 0; jj < nn; jj++)
                if (nbr[(size_t)jj * Np + nj] == i)
                    mj = (signed char)jj;
        }
        mirror[(size_t)j * Np + i] = mj;
    }
    nbi[i] = n;
    nb[i] = n;
}

// lpm_basic.c:77-89 / constitutive.c:546-558: first m < nb_initial with cs_initial[m] == -cs_initial[j] (per
// component within EPS); -1 if none.  Depends only on the initial geometry -> computed once.
__global__ void __launch_bounds__(BT)
derive_opposite_kernel(int N, int Np, const int *__restrict__ nbi, const double *__restrict__ cx0, const double *__restrict__ cy0,
                       const double *__restrict__ cz0, signed char *__restrict__ opp)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= N)
        return;
    const int n = nbi[i];
    for (int j = 0; j < n; j++) {
        const double ax = cx0[(size_t)j * Np + i], ay = cy0[(size_t)j * Np + i], az = cz0[(size_t)j * Np + i];
        signed char o = -1;
        for (int m = 0; m < n; m++) {
            if (fabs(cx0[(size_t)m * Np + i] + ax) < LPMB_EPS && fabs(cy0[(size_t)m * Np + i] + ay) < LPMB_EPS &&
                fabs(cz0[(size_t)m * Np + i] + az) < LPMB_EPS) {
                o = (signed char)m;
                break;
            }
        }
        opp[(size_t)j * Np + i] = o;
    }
}

// searchNormalNeighbor's geometric by-products (neighbor.c:23-26,34-37) from xyz for given lists
__global__ void __launch_bounds__(BT)
initial_geometry_kernel(int N, int Np, const int *__restrict__ nbr, const int *__restrict__ nbi, const double *__restrict__ xyz,
                        double *__restrict__ L0, double *__restrict__ cx0, double *__restrict__ cy0, double *__restrict__ cz0)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= N)
        return;
    const double xi = xyz[i], yi = xyz[(size_t)Np + i], zi = xyz[(size_t)2 * Np + i];
    const int n = nbi[i];
    for (int j = 0; j < n; j++) {
        const int nj = nbr[(size_t)j * Np + i];
        // dis = sqrt(pow(xj-xi,2)+pow(yj-yi,2)+pow(zj-zi,2)); cs = (xi-xj)/dis
        const double dx = xyz[nj] - xi, dy = xyz[(size_t)Np + nj] - yi, dz = xyz[(size_t)2 * Np + nj] - zi;
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        const size_t e = (size_t)j * Np + i;
        L0[e] = dis;
        cx0[e] = (xi - xyz[nj]) / dis;
        cy0[e] = (yi - xyz[(size_t)Np + nj]) / dis;
        cz0[e] = (zi - xyz[(size_t)2 * Np + nj]) / dis;
    }
}

int lpmb_derive_topology(lpmb_ctx *c, bool initial_geometry)
{
    int *nbr = fptr<int>(c, "neighbors");
    int *nbi = fptr<int>(c, "nb_initial"), *nb = fptr<int>(c, "nb");
    signed char *mirror = fptr<signed char>(c, "mirror"), *opp = fptr<signed char>(c, "oppslot");
    LPMB_REQUIRE(nbr && nbi && nb && mirror && opp, LPMB_ERR_STATE, "topology fields missing");
    const int g = lpmb_blocks(c->N, BT);
    derive_topology_kernel<<<g, BT, 0, c->stream>>>(c->N, c->Np, c->nn, nbr, nbi, nb, mirror);
    LPMB_LAUNCH_CHECK(c);
    double *L0 = fptr<double>(c, "distance_initial"), *cx0 = fptr<double>(c, "csx_initial"), *cy0 = fptr<double>(c, "csy_initial"),
           *cz0 = fptr<double>(c, "csz_initial");
    if (initial_geometry) {
        initial_geometry_kernel<<<g, BT, 0, c->stream>>>(c->N, c->Np, nbr, nbi, fptr<double>(c, "xyz"), L0, cx0, cy0, cz0);
        LPMB_LAUNCH_CHECK(c);
    }
    derive_opposite_kernel<<<g, BT, 0, c->stream>>>(c->N, c->Np, nbi, cx0, cy0, cz0, opp);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

extern "C" int lpmb_set_neighbors(lpmb_ctx *c, const int *neighbors, const int *nsign)
{
    LPMB_REQUIRE(c && neighbors && nsign, LPMB_ERR_ARG, "lpmb_set_neighbors: null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    const size_t cnt = (size_t)c->N * c->nn;
    LPMB_TRY(lpmb_field_set(c, "neighbors", neighbors, cnt));
    LPMB_TRY(lpmb_field_set(c, "nsign", nsign, cnt));
    LPMB_REQUIRE(lpmb_field(c, "xyz"), LPMB_ERR_STATE, "xyz missing");
    // distance_initial / cs*_initial are taken from the current xyz (= xyz_initial at set-up time) unless the
    // caller uploaded them already
    const bool have_geo = c->fields.count("distance_initial") && c->fields.count("csx_initial");
    return lpmb_derive_topology(c, !have_geo);
}

// ---------------------------------------------------------------------------------------------
// geometry pass  G(i; dLp*)   constitutive.c:241-260 / 495-515 / 625-645, lpm_basic.c:252-291
// ---------------------------------------------------------------------------------------------
// mode 0: dL = ((dis - L0) - dLp) * broken                  (all constitutive laws)
// mode 1: computedL(): dL = (dis - L0) - dLp, also writes distance[]
template <int MODE>
__global__ void __launch_bounds__(BT)
geometry_kernel(BondView v, const double *__restrict__ L0, const double *__restrict__ dLp, const double *__restrict__ broken,
                const double *__restrict__ Tv, double *__restrict__ dL, double *__restrict__ csx, double *__restrict__ csy,
                double *__restrict__ csz, double *__restrict__ dLt, double *__restrict__ TdLt, double *__restrict__ distance)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const double xi = v.xyz[i], yi = v.xyz[Np + i], zi = v.xyz[2 * Np + i];
    double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
    const int n = v.nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const double dx = xi - v.xyz[nj], dy = yi - v.xyz[Np + nj], dz = zi - v.xyz[2 * Np + nj];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        double d = dis - L0[e];
        d -= dLp[e];
        if (MODE == 0)
            d *= broken[e];
        else
            distance[e] = dis;
        dL[e] = d;
        const double td = Tv[e] * d;
        if (v.nsign[e] == 0) {
            t0 += d;
            T0 += td;
        } else {
            t1 += d;
            T1 += td;
        }
        csx[e] = dx / dis;
        csy[e] = dy / dis;
        csz[e] = dz / dis;
    }
    dLt[i] = t0;
    dLt[Np + i] = t1;
    TdLt[i] = T0;
    TdLt[Np + i] = T1;
}

// ---------------------------------------------------------------------------------------------
// bond force + internal force (owner pass)
// ---------------------------------------------------------------------------------------------
// LAW 6 (elastic, constitutive.c:267-279):  F = (2Kn dL + .5(TdLt_i+TdLt_j) + .5 Tv (dLt_i+dLt_j)) * broken
// LAW 0 (J2,      constitutive.c:652-667):  dL_ave = .5 (dL_ij + dL_ji);  F = (2Kn dL_ave + ...) * damage_w
// LAW 3 (J2 energy, constitutive.c:437-446): as LAW 0 but F *= (1 - damage_D[.][.][0]), which is not zero for a
//        broken bond -- and the reference only refreshes the dilatation sums of ii and its INTACT neighbours during
//        call ii (constitutive.c:289-295), so across a broken bond it reads whatever the partner holds at that point
//        of the serial particle loop: the new sums if an earlier call (the partner's own or one of its intact
//        neighbours', index < ii) already visited it, else the values from before computeBondForceGeneral.
//        `prev` holds those old sums so the serial result is reproduced exactly.
template <int LAW>
__global__ void __launch_bounds__(BT)
force_kernel(BondView v, const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ scale /* broken | w */,
             const double *__restrict__ dL, const double *__restrict__ dLt, const double *__restrict__ TdLt,
             const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ dL_ave,
             double *__restrict__ F, double *__restrict__ Pin, const double *__restrict__ broken = nullptr,
             const double *__restrict__ dLt_prev = nullptr, const double *__restrict__ TdLt_prev = nullptr)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const double dLt_i0 = dLt[i], dLt_i1 = dLt[Np + i], TdLt_i0 = TdLt[i], TdLt_i1 = TdLt[Np + i];
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    const int n = v.nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const int s = v.nsign[e];
        const double dLt_i = s ? dLt_i1 : dLt_i0, TdLt_i = s ? TdLt_i1 : TdLt_i0;
        double dLt_j = dLt[(size_t)s * Np + nj], TdLt_j = TdLt[(size_t)s * Np + nj];
        if (LAW == 3 && broken[e] <= LPMB_EPS) {
            // first call of the serial loop that refreshes nj: its own, or that of its lowest intact neighbour
            int first = nj;
            for (int m = 0; m < v.nn; m++) {
                const size_t em = (size_t)m * Np + nj;
                const int q = v.nbr[em];
                if (q != -1 && broken[em] > LPMB_EPS && q < first)
                    first = q;
            }
            if (first > i) {
                dLt_j = dLt_prev[(size_t)s * Np + nj];
                TdLt_j = TdLt_prev[(size_t)s * Np + nj];
            }
        }
        double stretch;
        if (LAW == 0 || LAW == 3) {
            const int mj = v.mirror[e];
            if (mj >= 0) {
                stretch = 0.5 * (dL[e] + dL[(size_t)mj * Np + nj]);
                dL_ave[e] = stretch;
            } else {
                stretch = dL_ave[e];
            }
        } else {
            stretch = dL[e];
        }
        double f = 2.0 * Kn[e] * stretch + 0.5 * (TdLt_i + TdLt_j) + 0.5 * Tv[e] * (dLt_i + dLt_j);
        if (LAW == 3)
            f *= (1.0 - scale[e]);
        else
            f *= scale[e];
        F[e] = f;
        p0 += csx[e] * f;
        p1 += csy[e] * f;
        p2 += csz[e] * f;
    }
    Pin[i] = p0;
    Pin[Np + i] = p1;
    Pin[2 * Np + i] = p2;
}

// ---------------------------------------------------------------------------------------------
// predictor (plmode 4)  constitutive.c:167-225
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
predictor_geometry_kernel(BondView v, const double *__restrict__ xyz_temp, const double *__restrict__ broken, const double *__restrict__ Tv,
                          double *__restrict__ ddL, double *__restrict__ ddLt, double *__restrict__ TddLt)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const double xi = v.xyz[i], yi = v.xyz[Np + i], zi = v.xyz[2 * Np + i];
    const double xt = xyz_temp[i], yt = xyz_temp[Np + i], zt = xyz_temp[2 * Np + i];
    double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
    const int n = v.nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const double ax = xt - xyz_temp[nj], ay = yt - xyz_temp[Np + nj], az = zt - xyz_temp[2 * Np + nj];
        const double dis0 = sqrt(ax * ax + ay * ay + az * az);
        const double bx = xi - v.xyz[nj], by = yi - v.xyz[Np + nj], bz = zi - v.xyz[2 * Np + nj];
        const double dis1 = sqrt(bx * bx + by * by + bz * bz);
        const double d = broken[e] * (dis1 - dis0);
        ddL[e] = d;
        const double td = Tv[e] * d;
        if (v.nsign[e] == 0) {
            t0 += d;
            T0 += td;
        } else {
            t1 += d;
            T1 += td;
        }
    }
    ddLt[i] = t0;
    ddLt[Np + i] = t1;
    TddLt[i] = T0;
    TddLt[Np + i] = T1;
}

__global__ void __launch_bounds__(BT)
predictor_force_kernel(BondView v, const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ broken,
                       const double *__restrict__ F_temp, const double *__restrict__ ddL, const double *__restrict__ ddLt,
                       const double *__restrict__ TddLt, const double *__restrict__ csx, const double *__restrict__ csy,
                       const double *__restrict__ csz, double *__restrict__ F, double *__restrict__ Pin)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    const int n = v.nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const int s = v.nsign[e];
        const size_t si = (size_t)s * Np + i, sj = (size_t)s * Np + nj;
        double f = F_temp[e] + 2.0 * Kn[e] * ddL[e] + 0.5 * (TddLt[si] + TddLt[sj]) + 0.5 * Tv[e] * (ddLt[si] + ddLt[sj]);
        f *= broken[e];
        F[e] = f;
        p0 += csx[e] * f;  // stale cs* on purpose (constitutive.c:197-199 are commented out)
        p1 += csy[e] * f;
        p2 += csz[e] * f;
    }
    Pin[i] = p0;
    Pin[Np + i] = p1;
    Pin[2 * Np + i] = p2;
}

// ---------------------------------------------------------------------------------------------
// opposite-bond factor (lpm_basic.c:72-90, constitutive.c:541-559)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double opp_flag(int nb_i, int nn, int o, const double *__restrict__ broken, size_t Np, int i)
{
    if (nb_i == nn)
        return 0.5;
    if (o < 0)
        return 1.0;
    return broken[(size_t)o * Np + i] <= LPMB_EPS ? 1.0 : 0.5;
}

// ---------------------------------------------------------------------------------------------
// J2 mixed linear hardening, return map per particle   constitutive.c:518-622
// reads slot-[0] state (dLp0, J2_beta0, J2_alpha0), writes slot-[2] (dLp2 = broken * (dLp0 + ddLp)),
// J2_beta2, J2_alpha2, J2_dlambda, ddLp, pl_flag
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
j2_return_map_kernel(BondView v, double V, double J2_H, double J2_xi, const double *__restrict__ Ce /* [ntype][3] */, const int *__restrict__ type,
                     const double *__restrict__ sigmay, const double *__restrict__ Kn, const double *__restrict__ Tv,
                     const double *__restrict__ w, const double *__restrict__ broken, const double *__restrict__ L0,
                     const double *__restrict__ dL, const double *__restrict__ dLt, const double *__restrict__ TdLt,
                     const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz,
                     const double *__restrict__ dLp0, const double *__restrict__ beta0, const double *__restrict__ alpha0,
                     double *__restrict__ dLp2, double *__restrict__ beta2, double *__restrict__ alpha2, double *__restrict__ ddLp,
                     double *__restrict__ dlambda, int *__restrict__ pl_flag)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i], nb_i = v.nb[i];
    const double dLt0 = dLt[i], dLt1 = dLt[Np + i], TdLt0 = TdLt[i], TdLt1 = TdLt[Np + i];
    double st[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int s = v.nsign[e];
        double Fij = 2.0 * Kn[e] * dL[e] + (s ? TdLt1 : TdLt0) + Tv[e] * (s ? dLt1 : dLt0);
        Fij *= w[e];
        const double of = opp_flag(nb_i, v.nn, v.opp[e], broken, Np, i);
        const double cx = csx[e], cy = csy[e], cz = csz[e];
        const double pre = of / V * L0[e] * Fij;
        st[0] += pre * cx * cx;
        st[1] += pre * cy * cy;
        st[2] += pre * cz * cz;
        st[3] += pre * cy * cz;
        st[4] += pre * cx * cz;
        st[5] += pre * cx * cy;
    }
    const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
    st[0] -= temp;
    st[1] -= temp;
    st[2] -= temp;
    double beta[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
        beta[q] = beta0[(size_t)q * Np + i];
        st[q] -= beta[q];
    }
    double seq = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        if (q < 3)
            seq += st[q] * st[q];
        else
            seq += 2.0 * st[q] * st[q];
    }
    seq = sqrt(3.0 / 2.0 * seq);
    double alpha = alpha0[i];
    double dl = 0.0;
    const double yield_func = seq - (sigmay[i] + (1.0 - J2_xi) * J2_H * alpha);
    if (yield_func > 0.0) {
        pl_flag[i] = 1;
        dl = yield_func / (3 * Ce[3 * type[i] + 2] + J2_H);
    }
    alpha += dl;
    double dpl[6] = {0, 0, 0, 0, 0, 0};
    if (fabs(seq) > LPMB_EPS) {
#pragma unroll
        for (int q = 0; q < 6; q++) {
            dpl[q] = dl * 1.5 * st[q] / seq;
            beta[q] += 2. / 3. * J2_xi * J2_H * dpl[q];
        }
    }
    for (int j = 0; j < v.nn; j++) {
        const size_t e = (size_t)j * Np + i;
        double xd = dLp0[e];
        if (j < n) {
            const double cx = csx[e], cy = csy[e], cz = csz[e];
            double dd = L0[e] * (dpl[0] * cx * cx + dpl[1] * cy * cy + dpl[2] * cz * cz + 2 * dpl[3] * cy * cz + 2 * dpl[4] * cx * cz +
                                 2 * dpl[5] * cx * cy);
            dd *= broken[e];
            ddLp[e] = dd;
            xd += dd;
        }
        dLp2[e] = broken[e] * xd;  // constitutive.c:670-671
    }
#pragma unroll
    for (int q = 0; q < 6; q++)
        beta2[(size_t)q * Np + i] = beta[q];
    alpha2[i] = alpha;
    dlambda[i] = dl;
}

// ---------------------------------------------------------------------------------------------
// Fused passes of the plmode-0 dispatcher (computeBondForceGeneral(0, t), constitutive.c:88-146 + 466-686 + lpm_basic.c:53-125).
// The five separate kernels re-read the per-bond arrays 3.7 times (78 GB of DRAM traffic for 21 GB of algorithmic bytes at
// 10 M particles, profiles/README.md).  Every pass is a pure per-particle function, so two kernels suffice:
//   j2_fused_kernel          = geometry(dLp[0]) -> return map -> geometry(dLp[2]): the first geometry stays in registers /
//                              local memory (its global outputs would be overwritten by the second one anyway),
//   j2_force_stress_kernel   = averaged bond force + Pin + computeStress (needs the neighbours' new dL / shell sums, hence
//                              the kernel boundary).
// Same expressions in the same order as the separate kernels (compiled with -fmad=false): bit-identical outputs, which the
// golden tests of the dispatcher check.  The separate kernels remain for plmode 1 / 3 / 6 and the per-particle entry points.
// ---------------------------------------------------------------------------------------------
#define J2F_MAXNN 18
__global__ void __launch_bounds__(BT)
j2_fused_kernel(BondView v, double V, double J2_H, double J2_xi, const double *__restrict__ Ce, const int *__restrict__ type,
                const double *__restrict__ sigmay, const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ w,
                const double *__restrict__ broken, const double *__restrict__ L0, const double *__restrict__ dLp0,
                const double *__restrict__ beta0, const double *__restrict__ alpha0, double *__restrict__ dLp2, double *__restrict__ beta2,
                double *__restrict__ alpha2, double *__restrict__ ddLp, double *__restrict__ dlambda, int *__restrict__ pl_flag,
                double *__restrict__ dL, double *__restrict__ csx, double *__restrict__ csy, double *__restrict__ csz, double *__restrict__ dLt,
                double *__restrict__ TdLt)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i], nb_i = v.nb[i];
    const double xi = v.xyz[i], yi = v.xyz[Np + i], zi = v.xyz[2 * Np + i];
    double dis_[J2F_MAXNN], d_[J2F_MAXNN], cx_[J2F_MAXNN], cy_[J2F_MAXNN], cz_[J2F_MAXNN];
    // ---- geometry with the slot-[0] plastic stretch (geometry_kernel<0>)
    double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const double dx = xi - v.xyz[nj], dy = yi - v.xyz[Np + nj], dz = zi - v.xyz[2 * Np + nj];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        double d = dis - L0[e];
        d -= dLp0[e];
        d *= broken[e];
        const double td = Tv[e] * d;
        if (v.nsign[e] == 0) {
            t0 += d;
            T0 += td;
        } else {
            t1 += d;
            T1 += td;
        }
        dis_[j] = dis;
        d_[j] = d;
        cx_[j] = dx / dis;
        cy_[j] = dy / dis;
        cz_[j] = dz / dis;
    }
    // ---- return map (j2_return_map_kernel)
    double st[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int s = v.nsign[e];
        double Fij = 2.0 * Kn[e] * d_[j] + (s ? T1 : T0) + Tv[e] * (s ? t1 : t0);
        Fij *= w[e];
        const double of = opp_flag(nb_i, v.nn, v.opp[e], broken, Np, i);
        const double cx = cx_[j], cy = cy_[j], cz = cz_[j];
        const double pre = of / V * L0[e] * Fij;
        st[0] += pre * cx * cx;
        st[1] += pre * cy * cy;
        st[2] += pre * cz * cz;
        st[3] += pre * cy * cz;
        st[4] += pre * cx * cz;
        st[5] += pre * cx * cy;
    }
    const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
    st[0] -= temp;
    st[1] -= temp;
    st[2] -= temp;
    double beta[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
        beta[q] = beta0[(size_t)q * Np + i];
        st[q] -= beta[q];
    }
    double seq = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        if (q < 3)
            seq += st[q] * st[q];
        else
            seq += 2.0 * st[q] * st[q];
    }
    seq = sqrt(3.0 / 2.0 * seq);
    double alpha = alpha0[i];
    double dl = 0.0;
    const double yield_func = seq - (sigmay[i] + (1.0 - J2_xi) * J2_H * alpha);
    if (yield_func > 0.0) {
        pl_flag[i] = 1;
        dl = yield_func / (3 * Ce[3 * type[i] + 2] + J2_H);
    }
    alpha += dl;
    double dpl[6] = {0, 0, 0, 0, 0, 0};
    if (fabs(seq) > LPMB_EPS) {
#pragma unroll
        for (int q = 0; q < 6; q++) {
            dpl[q] = dl * 1.5 * st[q] / seq;
            beta[q] += 2. / 3. * J2_xi * J2_H * dpl[q];
        }
    }
    // ---- new plastic stretch, then geometry with it (geometry_kernel<0> on slot [2]); unit vectors and distances are the same
    t0 = t1 = T0 = T1 = 0;
    for (int j = 0; j < v.nn; j++) {
        const size_t e = (size_t)j * Np + i;
        double xd = dLp0[e];
        const double b = broken[e];
        if (j < n) {
            const double cx = cx_[j], cy = cy_[j], cz = cz_[j];
            double dd = L0[e] * (dpl[0] * cx * cx + dpl[1] * cy * cy + dpl[2] * cz * cz + 2 * dpl[3] * cy * cz + 2 * dpl[4] * cx * cz +
                                 2 * dpl[5] * cx * cy);
            dd *= b;
            ddLp[e] = dd;
            xd += dd;
        }
        const double p2 = b * xd;  // constitutive.c:670-671
        dLp2[e] = p2;
        if (j < n) {
            double d = dis_[j] - L0[e];
            d -= p2;
            d *= b;
            dL[e] = d;
            const double td = Tv[e] * d;
            if (v.nsign[e] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
            csx[e] = cx_[j];
            csy[e] = cy_[j];
            csz[e] = cz_[j];
        }
    }
    dLt[i] = t0;
    dLt[Np + i] = t1;
    TdLt[i] = T0;
    TdLt[Np + i] = T1;
#pragma unroll
    for (int q = 0; q < 6; q++)
        beta2[(size_t)q * Np + i] = beta[q];
    alpha2[i] = alpha;
    dlambda[i] = dl;
}

// force_kernel<0> + stress_kernel in one pass over the bonds of i
__global__ void __launch_bounds__(BT)
j2_force_stress_kernel(BondView v, double V, const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ w,
                       const double *__restrict__ broken, const double *__restrict__ L0, const double *__restrict__ dL,
                       const double *__restrict__ dLt, const double *__restrict__ TdLt, const double *__restrict__ csx,
                       const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ dL_ave, double *__restrict__ F,
                       double *__restrict__ Pin, double *__restrict__ stress, double *__restrict__ seq_out, double *__restrict__ sm_out,
                       double *__restrict__ triax, double *__restrict__ bond_stress)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const double dLt_i0 = dLt[i], dLt_i1 = dLt[Np + i], TdLt_i0 = TdLt[i], TdLt_i1 = TdLt[Np + i];
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    double st[6] = {0, 0, 0, 0, 0, 0};
    const int n = v.nbi[i], nb_i = v.nb[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const int s = v.nsign[e];
        const double dLt_i = s ? dLt_i1 : dLt_i0, TdLt_i = s ? TdLt_i1 : TdLt_i0;
        const double dLt_j = dLt[(size_t)s * Np + nj], TdLt_j = TdLt[(size_t)s * Np + nj];
        double stretch;
        const int mj = v.mirror[e];
        if (mj >= 0) {
            stretch = 0.5 * (dL[e] + dL[(size_t)mj * Np + nj]);
            dL_ave[e] = stretch;
        } else {
            stretch = dL_ave[e];
        }
        double f = 2.0 * Kn[e] * stretch + 0.5 * (TdLt_i + TdLt_j) + 0.5 * Tv[e] * (dLt_i + dLt_j);
        f *= w[e];
        F[e] = f;
        const double cx = csx[e], cy = csy[e], cz = csz[e];
        p0 += cx * f;
        p1 += cy * f;
        p2 += cz * f;
        const double of = opp_flag(nb_i, v.nn, v.opp[e], broken, Np, i);
        const double pre = of / V * L0[e] * f;
        st[0] += pre * cx * cx;
        st[1] += pre * cy * cy;
        st[2] += pre * cz * cz;
        st[3] += pre * cy * cz;
        st[4] += pre * cx * cz;
        st[5] += pre * cx * cy;
    }
    Pin[i] = p0;
    Pin[Np + i] = p1;
    Pin[2 * Np + i] = p2;
    double seq = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        stress[(size_t)q * Np + i] = st[q];
        if (q < 3)
            seq += st[q] * st[q];
        else
            seq += 2.0 * st[q] * st[q];
    }
    seq = sqrt(3.0 / 2.0 * seq);
    const double sm = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
    seq_out[i] = seq;
    sm_out[i] = sm;
    if (seq > LPMB_EPS)
        triax[i] = sm / seq;  // keeps the previous value otherwise (lpm_basic.c:111-112)
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double cx = csx[e], cy = csy[e], cz = csz[e];
        bond_stress[e] = (st[0] * cx * cx + st[1] * cy * cy + st[2] * cz * cz + 2 * st[3] * cy * cz + 2 * st[4] * cx * cz + 2 * st[5] * cx * cy);
    }
}

// ---------------------------------------------------------------------------------------------
// J2 with the distortional-energy return map (plmode 3), per particle   constitutive.c:336-411
// Scalar equivalent back stress / plastic strain, plastic multiplier by bisection on (0,1) to TOLITER = 1e-4
// (14 halvings).  Quirks kept: the bond loops run over the first nb[i] slots (the CURRENT intact count, not
// nb_initial); J2_k is Kn of the last layer-1 bond among them; nb1 counts the layer-1 slots of the initial list.
// reads slot-[0] (dLp0, J2_beta_eq0, J2_alpha0), writes slot-[2], J2_dlambda, ddLp, pl_flag
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
j2_energy_return_map_kernel(BondView v, double V, double J2_H, double J2_xi, double radius, int load_indicator,
                            const double *__restrict__ Ce /* [ntype][3] */, const int *__restrict__ type, const double *__restrict__ sigmay,
                            const double *__restrict__ Kn, const double *__restrict__ broken, const double *__restrict__ dL,
                            const double *__restrict__ dLt, const double *__restrict__ dLp0, const double *__restrict__ beq0,
                            const double *__restrict__ alpha0, double *__restrict__ dLp2, double *__restrict__ beq2, double *__restrict__ alpha2,
                            double *__restrict__ ddLp, double *__restrict__ dlambda_out, int *__restrict__ pl_flag)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int nb_i = v.nb[i];
    int nb1 = 0;  // countNEqual(neighbors1[i], nneighbors1, -1): layer-1 entries of the initial list
    for (int j = 0; j < v.nbi[i]; j++)
        nb1 += v.nsign[(size_t)j * Np + i] == 0;
    const int nb2 = nb_i - nb1;
    const double dLt0 = dLt[i], dLt1 = dLt[Np + i];
    const double c44 = Ce[3 * type[i] + 2];
    double U_d = 0.0, J2_k = 0.0;
    const double J2_V = V * nb_i / v.nn;
    for (int j = 0; j < nb_i; j++) {
        const size_t e = (size_t)j * Np + i;
        const int s = v.nsign[e];
        if (s == 0) {
            U_d += 0.5 * Kn[e] * (dL[e] - dLt0 / nb1) * (dL[e] - dLt0 / nb1);
            J2_k = Kn[e];
        } else if (s == 1) {
            U_d += 0.5 * Kn[e] * (dL[e] - dLt1 / nb2) * (dL[e] - dLt1 / nb2);
        }
    }
    const double J2_sigma = sqrt(6.0 * c44 * U_d / J2_V);
    const double beq = beq0[i], alpha = alpha0[i], sy = sigmay[i];
    double dl = 0.0;
    double yield_func = fabs(load_indicator * J2_sigma - beq) - (sy + (1.0 - J2_xi) * J2_H * alpha);
    if (yield_func > 0.0) {
        pl_flag[i] = 1;
        double a = 0.0, b = 1.0;
        double ya = yield_func;
        while ((b - a) > 1e-4 /* TOLITER */) {
            dl = (a + b) / 2.0;
            yield_func = fabs(load_indicator * J2_sigma / (1.0 + 0.5 * dl) -
                              (beq + load_indicator * J2_xi * J2_H * dl * J2_sigma * sqrt(radius / J2_k / c44) / 6.0 / (1.0 + 0.5 * dl))) -
                         (sy + (1.0 - J2_xi) * J2_H * alpha + (1.0 - J2_xi) * J2_H * dl * J2_sigma * sqrt(radius / J2_k / c44) / 6.0 / (1.0 + 0.5 * dl));
            if (yield_func * ya < 0.0) {
                b = dl;
            } else {
                a = dl;
                ya = yield_func;
            }
        }
    }
    dlambda_out[i] = dl;
    alpha2[i] = alpha + dl / (1.0 + 0.5 * dl) * sqrt(radius / 6.0 / J2_k * U_d / J2_V);
    beq2[i] = (beq + load_indicator * J2_xi * J2_H * dl / (1.0 + 0.5 * dl) * sqrt(radius / 6.0 / J2_k * U_d / J2_V));
    double f_d = 0.0;
    for (int j = 0; j < v.nn; j++) {
        const size_t e = (size_t)j * Np + i;
        double xd = dLp0[e];
        if (j < nb_i) {
            const int s = v.nsign[e];
            if (s == 0)
                f_d = 2.0 * Kn[e] / (1 + dl / 2.0) * (dL[e] - dLt0 / nb1);
            if (s == 1)
                f_d = 2.0 * Kn[e] / (1 + dl / 2.0) * (dL[e] - dLt1 / nb2);
            double dd = dl * f_d / (4.0 * Kn[e]);
            dd *= broken[e];
            ddLp[e] = dd;
            xd += dd;
        }
        dLp2[e] = broken[e] * xd;  // constitutive.c:450-451
    }
}

// ---------------------------------------------------------------------------------------------
// J2 with nonlinear isotropic hardening (plmode 5)   constitutive.c:689-863
//
// The reference law is order dependent: call ii returns-maps ii AND each of its intact neighbours and writes the
// new plastic state straight into slot [0], so during one computeBondForceGeneral(5) particle i is updated once per
// "caller" c in C(i) = {i} u {intact neighbours of i}, in ascending c (the serial particle loop), each time starting
// from what the previous caller left.  Every one of those updates is a function of i's own state and of xyz only,
// so the whole chain of particle i can be run by one thread (trajectory kernel); what the callers need from it are
// snapshots: after the update made during call c, the elastic stretch of the bond towards c and i's dilatation sums
// (for c's own bond forces, constitutive.c:836-852).  The force kernel then forms F / Pin of ii from its own
// snapshot (taken at c = ii) and its neighbours' snapshots taken at c = ii.  Across a broken bond the partner is not
// refreshed by call ii: the reference reads what the partner's latest caller below ii left (or the values from
// before the call).  F[i] ends up holding the TRIAL forces (:737) of the last caller of i unless that is i itself.
// SY(x) = 620 + 3300 (1 - exp(-0.4 x)) (lpm.h:50); bisection on (0,1) to TOLITER = 1e-4.
// ---------------------------------------------------------------------------------------------
#define ISO_MAXNN 24

__global__ void __launch_bounds__(BT)
j2iso_trajectory_kernel(BondView v, double V, double J2_C, const double *__restrict__ Ce, const int *__restrict__ type,
                        const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ broken,
                        const double *__restrict__ L0, double *__restrict__ dLp0, double *__restrict__ alpha0, double *__restrict__ beta0,
                        double *__restrict__ dlambda_out, double *__restrict__ ddLp, double *__restrict__ dL, double *__restrict__ dLt,
                        double *__restrict__ TdLt, double *__restrict__ csx, double *__restrict__ csy, double *__restrict__ csz,
                        double *__restrict__ F, double *__restrict__ snap_dL /* [nn][Np] */, double *__restrict__ snap_t /* [nn][4][Np] */,
                        double *__restrict__ self_dL /* [nn][Np] */, double *__restrict__ self_t /* [4][Np] */, int *__restrict__ self_last,
                        int only_caller /* -1: the whole serial loop; ii: just call ii (computeBondForceJ2nonlinearIso(ii) on its own) */)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i], nb_i = v.nb[i];
    double gl[ISO_MAXNN], cx[ISO_MAXNN], cy[ISO_MAXNN], cz[ISO_MAXNN], dlp[ISO_MAXNN], d[ISO_MAXNN];
    const double xi = v.xyz[i], yi = v.xyz[Np + i], zi = v.xyz[2 * Np + i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const double dx = xi - v.xyz[nj], dy = yi - v.xyz[Np + nj], dz = zi - v.xyz[2 * Np + nj];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        gl[j] = dis - L0[e];
        cx[j] = dx / dis;
        cy[j] = dy / dis;
        cz[j] = dz / dis;
        dlp[j] = dLp0[e];
    }
    double alpha = alpha0[i], beta[6];
#pragma unroll
    for (int q = 0; q < 6; q++)
        beta[q] = beta0[(size_t)q * Np + i];
    const double c44 = Ce[3 * type[i] + 2];
    const double J2_V = V * nb_i / v.nn;
    double t0 = 0, t1 = 0, T0 = 0, T1 = 0, dl = 0.0;
    int last_is_self = 0;
    auto geometry = [&]() {
        t0 = t1 = T0 = T1 = 0.0;
        for (int j = 0; j < n; j++) {
            const size_t e = (size_t)j * Np + i;
            double x = gl[j];
            x -= dlp[j];
            x *= broken[e];
            d[j] = x;
            const double td = Tv[e] * x;
            if (v.nsign[e] == 0) {
                t0 += x;
                T0 += td;
            } else {
                t1 += x;
                T1 += td;
            }
        }
    };
    // callers in ascending particle index: neighbour slots are ascending (neighbor.c:29,40), i itself slots in between
    bool self_done = false, ran = false;
    for (int s = 0; s <= n; s++) {
        int caller_slot;  // -1 = i itself
        if (!self_done && (s == n || v.nbr[(size_t)s * Np + i] > i)) {
            caller_slot = -1;
            self_done = true;
            s--;  // the slot is looked at again after the self call
        } else {
            if (s == n)
                break;
            caller_slot = s;
            if (!(broken[(size_t)s * Np + i] > LPMB_EPS))
                continue;  // not in that particle's list (constitutive.c:694-699)
        }
        if (only_caller >= 0 && (caller_slot < 0 ? i : v.nbr[(size_t)caller_slot * Np + i]) != only_caller)
            continue;
        ran = true;
        // (A) elastic stretches with the current plastic stretch   :703-721
        geometry();
        // (B) trial force / stress, return map   :729-809
        double st[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 0; j < n; j++) {
            const size_t e = (size_t)j * Np + i;
            const int sg = v.nsign[e];
            double f = 2.0 * Kn[e] * d[j] + (sg ? T1 : T0) + Tv[e] * (sg ? t1 : t0);
            f *= broken[e];
            F[e] = f;
            const double pre = 0.5 / J2_V * L0[e] * f;
            st[0] += pre * cx[j] * cx[j];
            st[1] += pre * cy[j] * cy[j];
            st[2] += pre * cz[j] * cz[j];
            st[3] += pre * cy[j] * cz[j];
            st[4] += pre * cx[j] * cz[j];
            st[5] += pre * cx[j] * cy[j];
        }
        const double temp = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
        st[0] -= temp;
        st[1] -= temp;
        st[2] -= temp;
#pragma unroll
        for (int q = 0; q < 6; q++)
            st[q] -= beta[q];
        double seq = 0.0;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            if (q < 3)
                seq += st[q] * st[q];
            else
                seq += 2.0 * st[q] * st[q];
        }
        seq = sqrt(3.0 / 2.0 * seq);
        dl = 0.0;
        double yield_func = seq - (620.0 + 3300.0 * (1.0 - exp(-0.4 * alpha)));
        if (yield_func > 0.0) {
            double a = 0.0, b = 1.0, ya = yield_func;
            while ((b - a) > 1e-4 /* TOLITER */) {
                dl = (a + b) / 2.0;
                // SY(J2_alpha + dlambda) expands, unparenthesised (lpm.h:50), to exp(-0.4 * J2_alpha + dlambda): kept
                yield_func = seq - 1.5 * dl * (2.0 * c44 + J2_C) - (620.0 + 3300.0 * (1.0 - exp(-0.4 * alpha + dl)));
                if (yield_func * ya < 0.0) {
                    b = dl;
                } else {
                    a = dl;
                    ya = yield_func;
                }
            }
        }
        alpha += dl;
        double dpl[6];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            dpl[q] = dl * 1.5 * st[q] / seq;  // no guard for seq == 0 in the reference either (:794)
            beta[q] += J2_C * dpl[q];
        }
        for (int j = 0; j < n; j++) {
            const size_t e = (size_t)j * Np + i;
            const double dd = L0[e] * (dpl[0] * cx[j] * cx[j] + dpl[1] * cy[j] * cy[j] + dpl[2] * cz[j] * cz[j] + 2 * dpl[3] * cy[j] * cz[j] +
                                       2 * dpl[4] * cx[j] * cz[j] + 2 * dpl[5] * cx[j] * cy[j]);
            ddLp[e] = dd;
            dlp[j] += dd;
        }
        // (C) elastic stretches with the new plastic stretch   :815-831
        geometry();
        // what the caller reads in its own force pass
        if (caller_slot < 0) {
            for (int j = 0; j < n; j++)
                self_dL[(size_t)j * Np + i] = d[j];
            self_t[i] = t0;
            self_t[Np + i] = t1;
            self_t[2 * Np + i] = T0;
            self_t[3 * Np + i] = T1;
            last_is_self = 1;
        } else {
            snap_dL[(size_t)caller_slot * Np + i] = d[caller_slot];
            double *q = snap_t + (size_t)caller_slot * 4 * Np + i;
            q[0] = t0;
            q[Np] = t1;
            q[2 * Np] = T0;
            q[3 * Np] = T1;
            last_is_self = 0;
        }
    }
    if (only_caller >= 0 && !ran)
        return;  // not in the star of the one call that is made: the reference does not touch this particle
    // what the serial loop leaves behind
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        dLp0[e] = dlp[j];
        dL[e] = d[j];
        csx[e] = cx[j];
        csy[e] = cy[j];
        csz[e] = cz[j];
    }
    dLt[i] = t0;
    dLt[Np + i] = t1;
    TdLt[i] = T0;
    TdLt[Np + i] = T1;
    alpha0[i] = alpha;
#pragma unroll
    for (int q = 0; q < 6; q++)
        beta0[(size_t)q * Np + i] = beta[q];
    dlambda_out[i] = dl;
    self_last[i] = last_is_self;
}

// bond forces of ii at the moment of its own call   constitutive.c:833-857
__global__ void __launch_bounds__(BT)
j2iso_force_kernel(BondView v, const double *__restrict__ Kn, const double *__restrict__ Tv, const double *__restrict__ w,
                   const double *__restrict__ broken, const double *__restrict__ snap_dL, const double *__restrict__ snap_t,
                   const double *__restrict__ self_dL, const double *__restrict__ self_t, const int *__restrict__ self_last,
                   const double *__restrict__ dL_prev, const double *__restrict__ dLt_prev, const double *__restrict__ TdLt_prev,
                   const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ dL_ave,
                   double *__restrict__ F, double *__restrict__ Pin,
                   int only_row /* -1: every particle; ii: the force pass of the single call ii (partners across broken bonds are then
                                   read as they are in memory = the `prev` arrays, which the caller points at the live fields) */)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N || (only_row >= 0 && i != only_row))
        return;
    const size_t Np = v.Np;
    const double ti[4] = {self_t[i], self_t[Np + i], self_t[2 * Np + i], self_t[3 * Np + i]};
    const bool write_F = self_last[i] != 0;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    const int n = v.nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = v.nbr[e];
        const int s = v.nsign[e];
        const int mj = v.mirror[e];
        double dL_j = 0.0, dLt_j, TdLt_j;
        if (broken[e] > LPMB_EPS && mj >= 0) {
            // nj was refreshed during this very call: its snapshot for caller i sits at its slot towards i
            const double *q = snap_t + (size_t)mj * 4 * Np + nj;
            dL_j = snap_dL[(size_t)mj * Np + nj];
            dLt_j = q[(size_t)s * Np];
            TdLt_j = q[(size_t)(2 + s) * Np];
        } else {
            // broken bond: nj holds what its latest caller below i left, or the values from before the whole call
            int best = -2, best_slot = -1;  // -2: none yet; slot -1 = nj itself
            if (only_row < 0 && nj < i)
                best = nj;
            for (int m = 0; only_row < 0 && m < v.nn; m++) {
                const size_t em = (size_t)m * Np + nj;
                const int q = v.nbr[em];
                if (q != -1 && broken[em] > LPMB_EPS && q < i && q > best) {
                    best = q;
                    best_slot = m;
                }
            }
            if (best == -2) {
                dLt_j = dLt_prev[(size_t)s * Np + nj];
                TdLt_j = TdLt_prev[(size_t)s * Np + nj];
                if (mj >= 0)
                    dL_j = dL_prev[(size_t)mj * Np + nj];
            } else if (best_slot < 0) {
                dLt_j = self_t[(size_t)s * Np + nj];
                TdLt_j = self_t[(size_t)(2 + s) * Np + nj];
                if (mj >= 0)
                    dL_j = self_dL[(size_t)mj * Np + nj];
            } else {
                const double *q = snap_t + (size_t)best_slot * 4 * Np + nj;
                dLt_j = q[(size_t)s * Np];
                TdLt_j = q[(size_t)(2 + s) * Np];
                dL_j = 0.0;  // the stretch of a broken bond is 0 in every snapshot (x broken)
            }
        }
        const double dL_i = self_dL[e];
        double stretch;
        if (mj >= 0) {
            stretch = 0.5 * (dL_i + dL_j);
            dL_ave[e] = stretch;
        } else {
            stretch = dL_ave[e];
        }
        double f = 2.0 * Kn[e] * stretch + 0.5 * (ti[2 + s] + TdLt_j) + 0.5 * Tv[e] * (ti[s] + dLt_j);
        f *= w[e];
        if (write_F)
            F[e] = f;
        p0 += csx[e] * f;
        p1 += csy[e] * f;
        p2 += csz[e] * f;
    }
    Pin[i] = p0;
    Pin[Np + i] = p1;
    Pin[2 * Np + i] = p2;
}

// ---------------------------------------------------------------------------------------------
// stress   lpm_basic.c:53-125
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
stress_kernel(BondView v, double V, const double *__restrict__ L0, const double *__restrict__ F, const double *__restrict__ broken,
              const double *__restrict__ csx, const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ stress,
              double *__restrict__ seq_out, double *__restrict__ sm_out, double *__restrict__ triax, double *__restrict__ bond_stress)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i], nb_i = v.nb[i];
    double st[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double of = opp_flag(nb_i, v.nn, v.opp[e], broken, Np, i);
        const double cx = csx[e], cy = csy[e], cz = csz[e];
        const double pre = of / V * L0[e] * F[e];
        st[0] += pre * cx * cx;
        st[1] += pre * cy * cy;
        st[2] += pre * cz * cz;
        st[3] += pre * cy * cz;
        st[4] += pre * cx * cz;
        st[5] += pre * cx * cy;
    }
    double seq = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        stress[(size_t)q * Np + i] = st[q];
        if (q < 3)
            seq += st[q] * st[q];
        else
            seq += 2.0 * st[q] * st[q];
    }
    seq = sqrt(3.0 / 2.0 * seq);
    const double sm = 1.0 / 3.0 * (st[0] + st[1] + st[2]);
    seq_out[i] = seq;
    sm_out[i] = sm;
    if (seq > LPMB_EPS)
        triax[i] = sm / seq;  // keeps the previous value otherwise (lpm_basic.c:111-112)
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double cx = csx[e], cy = csy[e], cz = csz[e];
        bond_stress[e] = (st[0] * cx * cx + st[1] * cy * cy + st[2] * cz * cz + 2 * st[3] * cy * cz + 2 * st[4] * cx * cz + 2 * st[5] * cx * cy);
    }
}

// ---------------------------------------------------------------------------------------------
// updateCrack   constitutive.c:1399-1434
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT)
update_crack_kernel(BondView v, const double *__restrict__ broken, const double *__restrict__ w, const double *__restrict__ csx,
                    const double *__restrict__ csy, const double *__restrict__ csz, double *__restrict__ F, double *__restrict__ Pin,
                    double *__restrict__ damage_visual, int *__restrict__ fix_index)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i];
    int nbv = n;
    double vis = 0.0, p0 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        if (broken[e] <= LPMB_EPS)
            nbv -= 1;
        vis += broken[e];
        const double f = F[e] * w[e];  // scaled again on purpose
        F[e] = f;
        p0 += csx[e] * f;
        p1 += csy[e] * f;
        p2 += csz[e] * f;
    }
    v.nb[i] = nbv;
    Pin[i] = p0;
    Pin[Np + i] = p1;
    Pin[2 * Np + i] = p2;
    if (nbv < 1)
        for (int k = 0; k < v.dim; k++)
            fix_index[(size_t)k * Np + i] = 0;
    damage_visual[i] = 1 - vis / n;
}

// ---------------------------------------------------------------------------------------------
// updateRR + norms   stiffness.c:519-534, lpmc_project.c:412-413
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// residual[k] = dispBC_index[k] * (Pex[k] - Pin[k]); partial sums of residual^2 and of Pin^2 over constrained DoFs
__global__ void __launch_bounds__(256)
update_rr_kernel(int dim, int Np, int N, int own0, int own1, const int *__restrict__ bc, const double *__restrict__ Pex,
                 const double *__restrict__ Pin, double *__restrict__ residual, double *__restrict__ partials /* [2][grid] */)
{
    __shared__ double red[2][8];
    double s_res = 0.0, s_rea = 0.0;
    const size_t n = (size_t)dim * Np;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) {
        const int i = (int)(e % Np);
        if (i >= N)
            continue;
        const int b = bc[e];
        const double pin = Pin[e];  // Pin is [3][Np]; the first dim components line up with the DoF layout
        const double r = b * (Pex[e] - pin);
        residual[e] = r;
        if (i < own0 || i >= own1)
            continue;  // ghost rows (multi-GPU) belong to the neighbouring slab's norm
        s_res += r * r;
        if (b == 0)
            s_rea += pin * pin;
    }
    s_res = warp_sum_d(s_res);
    s_rea = warp_sum_d(s_rea);
    const int wdx = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[0][wdx] = s_res;
        red[1][wdx] = s_rea;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b2 = 0;
        for (int k = 0; k < 8; k++) {
            a += red[0][k];
            b2 += red[1][k];
        }
        partials[blockIdx.x] = a;
        partials[gridDim.x + blockIdx.x] = b2;
    }
}

__global__ void finish_rr_kernel(const double *__restrict__ partials, int nparts, double *__restrict__ out2)
{
    // single block, fixed order
    __shared__ double red[2][8];
    double a = 0, b = 0;
    for (int k = threadIdx.x; k < nparts; k += 256) {
        a += partials[k];
        b += partials[nparts + k];
    }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    const int wdx = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[0][wdx] = a;
        red[1][wdx] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0;
        for (int k = 0; k < 8; k++) {
            x += red[0][k];
            y += red[1][k];
        }
        out2[0] = x;  // squared norms; the square root is taken after the (optional) all-reduce
        out2[1] = y;
    }
}

// ---------------------------------------------------------------------------------------------
// strain   lpm_basic.c:127-249: per particle, weighted least squares over the initial bond directions
// (weights 0.1 / 0.9 for the two shells), n x n system (n = 3(dim-1)) solved by LU with partial pivoting
// (LAPACKE_dgesv, row-major; first row of maximal |a| wins).  Singular system (info > 0): dgesv leaves the right
// hand side untouched and the reference then stores exactly that (its copy from particle i-1 is overwritten in
// 3-D; in 2-D it only reaches component [2], which nothing else ever writes); particle 0 is skipped.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(BT)
strain_kernel(BondView v, const double *__restrict__ xyz0, const double *__restrict__ L0, const double *__restrict__ dL,
              double *__restrict__ strain)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    constexpr int NS = 3 * (DIM - 1);
    const size_t Np = v.Np;
    double r4[DIM][DIM][DIM][DIM], r2[DIM][DIM];
    for (int k = 0; k < DIM; k++)
        for (int n = 0; n < DIM; n++) {
            r2[k][n] = 0.0;
            for (int m = 0; m < DIM; m++)
                for (int l = 0; l < DIM; l++)
                    r4[k][n][m][l] = 0.0;
        }
    double xi[DIM];
    for (int k = 0; k < DIM; k++)
        xi[k] = xyz0[(size_t)k * Np + i];
    const int nbi = v.nbi[i];
    for (int j = 0; j < nbi; j++) {
        const size_t e = (size_t)j * Np + i;
        const int s = v.nsign[e];
        if (s != 0 && s != 1)
            continue;
        const double w = s == 0 ? 0.1 : 0.9;
        const int nj = v.nbr[e];
        const double L = L0[e], d = dL[e];
        double dx[DIM];
        for (int k = 0; k < DIM; k++)
            dx[k] = xyz0[(size_t)k * Np + nj] - xi[k];
        for (int k = 0; k < DIM; k++)
            for (int n = 0; n < DIM; n++) {
                for (int m = 0; m < DIM; m++)
                    for (int l = 0; l < DIM; l++)
                        r4[k][n][m][l] += w * dx[k] / L * dx[n] / L * dx[m] / L * dx[l] / L;
                r2[k][n] += w * d / L * dx[k] / L * dx[n] / L;
            }
    }
    double a[NS * NS], b[NS], b_in[NS];
    int ii = 0;
    for (int j = 0; j < DIM; j++)
        for (int k = 0; k < DIM; k++)
            for (int m = 0; m < DIM; m++)
                for (int l = 0; l < DIM; l++)
                    if (m <= l && j <= k)
                        a[ii++] = r4[m][l][j][k];
    ii = 0;
    for (int j = 0; j < DIM; j++)
        for (int k = 0; k < DIM; k++)
            if (j <= k) {
                b[ii] = r2[j][k];
                b_in[ii] = r2[j][k];
                ii++;
            }
    int info = 0;
    for (int k = 0; k < NS; k++) {
        int p = k;
        double amax = fabs(a[k * NS + k]);
        for (int r = k + 1; r < NS; r++) {
            const double t = fabs(a[r * NS + k]);
            if (t > amax) {
                amax = t;
                p = r;
            }
        }
        if (a[p * NS + k] == 0.0) {
            if (info == 0)
                info = k + 1;
            continue;
        }
        if (p != k) {
            for (int c = 0; c < NS; c++) {
                const double t = a[k * NS + c];
                a[k * NS + c] = a[p * NS + c];
                a[p * NS + c] = t;
            }
            const double t = b[k];
            b[k] = b[p];
            b[p] = t;
        }
        const double piv = a[k * NS + k];
        for (int r = k + 1; r < NS; r++) {
            const double l = a[r * NS + k] / piv;
            a[r * NS + k] = l;
            if (l != 0.0) {
                for (int c = k + 1; c < NS; c++)
                    a[r * NS + c] -= l * a[k * NS + c];
                b[r] -= l * b[k];
            }
        }
    }
    if (info != 0) {
        if (i == 0)
            return;
        for (int k = 0; k < NS; k++)
            b[k] = b_in[k];
    } else {
        for (int r = NS - 1; r >= 0; r--) {
            double t = b[r];
            for (int c = r + 1; c < NS; c++)
                t -= a[r * NS + c] * b[c];
            b[r] = t / a[r * NS + r];
        }
    }
    if (DIM == 2) {
        strain[i] = b[0];
        strain[5 * Np + i] = b[1];
        strain[Np + i] = b[2];
    } else {
        strain[i] = b[0];
        strain[5 * Np + i] = b[1];
        strain[4 * Np + i] = b[2];
        strain[Np + i] = b[3];
        strain[3 * Np + i] = b[4];
        strain[2 * Np + i] = b[5];
    }
}

extern "C" int lpmb_compute_strain(lpmb_ctx *c)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    BondView v;
    LPMB_TRY(make_view(c, v));
    const double *x0 = fptr<double>(c, "xyz_initial"), *L0 = fptr<double>(c, "distance_initial"), *dL = fptr<double>(c, "dL");
    double *strain = fptr<double>(c, "strain_tensor");
    LPMB_REQUIRE(x0 && L0 && dL && strain, LPMB_ERR_STATE, "computeStrain: fields missing");
    if (c->dim == 3)
        strain_kernel<3><<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(v, x0, L0, dL, strain);
    else
        strain_kernel<2><<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(v, x0, L0, dL, strain);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

// ---------------------------------------------------------------------------------------------
// host-side orchestration
// ---------------------------------------------------------------------------------------------
static int copy_field(lpmb_ctx *c, const char *dst, const char *src)
{
    Field *d = lpmb_field(c, dst), *s = lpmb_field(c, src);
    LPMB_REQUIRE(d && s && d->count == s->count && d->elem() == s->elem(), LPMB_ERR_STATE, "copy_field %s <- %s: mismatch", dst, src);
    LPMB_CUDA(cudaMemcpyAsync(d->d, s->d, d->count * d->elem(), cudaMemcpyDeviceToDevice, c->stream));
    return LPMB_OK;
}

// switchStateV   constitutive.c:10-85.  (cp_* state joins when the crystal-plasticity law is built.)
extern "C" int lpmb_switch_state(lpmb_ctx *c, int flag)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    if (flag == 0) {  // [0] := [1]
        LPMB_TRY(copy_field(c, "dLp0", "dLp1"));
        LPMB_TRY(copy_field(c, "damage_D0", "damage_D1"));
        LPMB_TRY(copy_field(c, "J2_beta0", "J2_beta1"));
        LPMB_TRY(copy_field(c, "J2_alpha0", "J2_alpha1"));
        LPMB_TRY(copy_field(c, "J2_beta_eq0", "J2_beta_eq1"));
        LPMB_TRY(copy_field(c, "damage_local0", "damage_local1"));
        LPMB_TRY(copy_field(c, "damage_nonlocal0", "damage_nonlocal1"));
    } else if (flag == 1) {  // [1] := [0]
        LPMB_TRY(copy_field(c, "dLp1", "dLp0"));
        LPMB_TRY(copy_field(c, "damage_D1", "damage_D0"));
        LPMB_TRY(copy_field(c, "J2_beta1", "J2_beta0"));
        LPMB_TRY(copy_field(c, "J2_alpha1", "J2_alpha0"));
        LPMB_TRY(copy_field(c, "J2_beta_eq1", "J2_beta_eq0"));
        LPMB_TRY(copy_field(c, "damage_local1", "damage_local0"));
        LPMB_TRY(copy_field(c, "damage_nonlocal1", "damage_nonlocal0"));
    } else if (flag == 2) {  // [0] := [2] (no damage_* here, constitutive.c:63-84)
        LPMB_TRY(copy_field(c, "dLp0", "dLp2"));
        LPMB_TRY(copy_field(c, "J2_beta0", "J2_beta2"));
        LPMB_TRY(copy_field(c, "J2_alpha0", "J2_alpha2"));
        LPMB_TRY(copy_field(c, "J2_beta_eq0", "J2_beta_eq2"));
    } else {
        lpmb_set_error("switchStateV: flag %d", flag);
        return LPMB_ERR_ARG;
    }
    // crystal-plasticity state joins when slip systems are defined (nslipSys > 0, constitutive.c:25-35,50-60,74-82)
    if (param(c, "nslipSys", 0.0) > 0.0) {
        static const char *dst[3] = {"0", "1", "0"}, *src[3] = {"1", "0", "2"};
        for (const char *base : {"cp_gy", "cp_A_single", "cp_A"}) {
            const std::string d = std::string(base) + dst[flag], s2 = std::string(base) + src[flag];
            LPMB_TRY(copy_field(c, d.c_str(), s2.c_str()));
        }
    }
    return LPMB_OK;
}

extern "C" int lpmb_compute_dl(lpmb_ctx *c)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    BondView v;
    LPMB_TRY(make_view(c, v));
    geometry_kernel<1><<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(
        v, fptr<double>(c, "distance_initial"), fptr<double>(c, "dLp0"), fptr<double>(c, "damage_broken"), fptr<double>(c, "Tv"),
        fptr<double>(c, "dL"), fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"), fptr<double>(c, "dL_total"),
        fptr<double>(c, "TdL_total"), fptr<double>(c, "distance"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

static int run_geometry(lpmb_ctx *c, BondView &v, const char *dLp_slot)
{
    geometry_kernel<0><<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(
        v, fptr<double>(c, "distance_initial"), fptr<double>(c, dLp_slot), fptr<double>(c, "damage_broken"), fptr<double>(c, "Tv"),
        fptr<double>(c, "dL"), fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"), fptr<double>(c, "dL_total"),
        fptr<double>(c, "TdL_total"), nullptr);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

int lpmb_compute_stress(lpmb_ctx *c)
{
    BondView v;
    LPMB_TRY(make_view(c, v));
    stress_kernel<<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(
        v, param(c, "particle_volume"), fptr<double>(c, "distance_initial"), fptr<double>(c, "F"), fptr<double>(c, "damage_broken"),
        fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"), fptr<double>(c, "stress_tensor"),
        fptr<double>(c, "J2_stresseq"), fptr<double>(c, "J2_stressm"), fptr<double>(c, "J2_triaxiality"), fptr<double>(c, "bond_stress"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

// computeBondForceGeneral(plmode, t)   constitutive.c:88-146
// scratch of the plmode-5 law: per-caller snapshots, the particle's own snapshot, the fields as they were before the call
static int iso_scratch(lpmb_ctx *c)
{
    if (!c->fields.count("iso_snap_dL")) {
        LPMB_TRY(lpmb_field_alloc(c, "iso_snap_dL", FK_BOND, FT_F64, c->nn));
        LPMB_TRY(lpmb_field_alloc(c, "iso_snap_t", FK_PART, FT_F64, 4 * c->nn));
        LPMB_TRY(lpmb_field_alloc(c, "iso_self_dL", FK_BOND, FT_F64, c->nn));
        LPMB_TRY(lpmb_field_alloc(c, "iso_self_t", FK_PART, FT_F64, 4));
        LPMB_TRY(lpmb_field_alloc(c, "iso_self_last", FK_PART, FT_I32, 1));
        LPMB_TRY(lpmb_field_alloc(c, "dL_prev", FK_BOND, FT_F64, c->nn));
    }
    if (!c->fields.count("dL_total_prev")) {
        LPMB_TRY(lpmb_field_alloc(c, "dL_total_prev", FK_PART, FT_F64, 2));
        LPMB_TRY(lpmb_field_alloc(c, "TdL_total_prev", FK_PART, FT_F64, 2));
    }
    return LPMB_OK;
}

extern "C" int lpmb_bond_force(lpmb_ctx *c, int plmode, int load_indicator)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_REQUIRE(c->params.count("particle_volume"), LPMB_ERR_STATE, "parameter particle_volume not set");
    BondView v;
    LPMB_TRY(make_view(c, v));
    const int g = lpmb_blocks(c->N, BT);
    double *Kn = fptr<double>(c, "Kn"), *Tv = fptr<double>(c, "Tv"), *broken = fptr<double>(c, "damage_broken"), *w = fptr<double>(c, "damage_w");
    double *dL = fptr<double>(c, "dL"), *dLt = fptr<double>(c, "dL_total"), *TdLt = fptr<double>(c, "TdL_total");
    double *csx = fptr<double>(c, "csx"), *csy = fptr<double>(c, "csy"), *csz = fptr<double>(c, "csz");
    double *F = fptr<double>(c, "F"), *Pin = fptr<double>(c, "Pin"), *dL_ave = fptr<double>(c, "dL_ave");
    if (plmode == 6) {
        LPMB_TRY(run_geometry(c, v, "dLp0"));
        force_kernel<6><<<g, BT, 0, c->stream>>>(v, Kn, Tv, broken, dL, dLt, TdLt, csx, csy, csz, dL_ave, F, Pin);
        LPMB_LAUNCH_CHECK(c);
    } else if (plmode == 4) {
        double *ddL = fptr<double>(c, "ddL"), *ddLt = fptr<double>(c, "ddL_total"), *TddLt = fptr<double>(c, "TddL_total");
        predictor_geometry_kernel<<<g, BT, 0, c->stream>>>(v, fptr<double>(c, "xyz_temp"), broken, Tv, ddL, ddLt, TddLt);
        LPMB_LAUNCH_CHECK(c);
        predictor_force_kernel<<<g, BT, 0, c->stream>>>(v, Kn, Tv, broken, fptr<double>(c, "F_temp"), ddL, ddLt, TddLt, csx, csy, csz, F, Pin);
        LPMB_LAUNCH_CHECK(c);
    } else if (plmode == 0) {
        LPMB_REQUIRE(c->params.count("J2_H") && c->params.count("J2_xi"), LPMB_ERR_STATE, "J2_H / J2_xi not set");
        Field *ce = lpmb_field(c, "Ce");
        LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
        if (c->nn <= J2F_MAXNN && param(c, "j2_fused", 1.0) != 0.0) {
            // two fused passes instead of five (same arithmetic, bit-identical outputs; see j2_fused_kernel)
            j2_fused_kernel<<<g, BT, 0, c->stream>>>(
                v, param(c, "particle_volume"), param(c, "J2_H"), param(c, "J2_xi"), (const double *)ce->d, fptr<int>(c, "type"),
                fptr<double>(c, "sigmay"), Kn, Tv, w, broken, fptr<double>(c, "distance_initial"), fptr<double>(c, "dLp0"),
                fptr<double>(c, "J2_beta0"), fptr<double>(c, "J2_alpha0"), fptr<double>(c, "dLp2"), fptr<double>(c, "J2_beta2"),
                fptr<double>(c, "J2_alpha2"), fptr<double>(c, "ddLp"), fptr<double>(c, "J2_dlambda"), fptr<int>(c, "pl_flag"), dL, csx, csy, csz,
                dLt, TdLt);
            LPMB_LAUNCH_CHECK(c);
            j2_force_stress_kernel<<<g, BT, 0, c->stream>>>(
                v, param(c, "particle_volume"), Kn, Tv, w, broken, fptr<double>(c, "distance_initial"), dL, dLt, TdLt, csx, csy, csz, dL_ave, F,
                Pin, fptr<double>(c, "stress_tensor"), fptr<double>(c, "J2_stresseq"), fptr<double>(c, "J2_stressm"),
                fptr<double>(c, "J2_triaxiality"), fptr<double>(c, "bond_stress"));
            LPMB_LAUNCH_CHECK(c);
            return lpmb_switch_state(c, 2);
        }
        LPMB_TRY(run_geometry(c, v, "dLp0"));
        j2_return_map_kernel<<<g, BT, 0, c->stream>>>(
            v, param(c, "particle_volume"), param(c, "J2_H"), param(c, "J2_xi"), (const double *)ce->d, fptr<int>(c, "type"),
            fptr<double>(c, "sigmay"), Kn, Tv, w, broken, fptr<double>(c, "distance_initial"), dL, dLt, TdLt, csx, csy, csz,
            fptr<double>(c, "dLp0"), fptr<double>(c, "J2_beta0"), fptr<double>(c, "J2_alpha0"), fptr<double>(c, "dLp2"),
            fptr<double>(c, "J2_beta2"), fptr<double>(c, "J2_alpha2"), fptr<double>(c, "ddLp"), fptr<double>(c, "J2_dlambda"),
            fptr<int>(c, "pl_flag"));
        LPMB_LAUNCH_CHECK(c);
        LPMB_TRY(run_geometry(c, v, "dLp2"));
        force_kernel<0><<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, dL, dLt, TdLt, csx, csy, csz, dL_ave, F, Pin);
        LPMB_LAUNCH_CHECK(c);
    } else if (plmode == 3) {
        // J2, distortional-energy return map (constitutive.c:286-463): geometry -> return map -> geometry -> force
        LPMB_REQUIRE(c->params.count("J2_H") && c->params.count("J2_xi") && c->params.count("radius"), LPMB_ERR_STATE, "J2_H / J2_xi / radius not set");
        Field *ce = lpmb_field(c, "Ce");
        LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
        // dilatation sums as they are before this call (see force_kernel, LAW 3)
        if (!c->fields.count("dL_total_prev")) {
            LPMB_TRY(lpmb_field_alloc(c, "dL_total_prev", FK_PART, FT_F64, 2));
            LPMB_TRY(lpmb_field_alloc(c, "TdL_total_prev", FK_PART, FT_F64, 2));
        }
        LPMB_TRY(copy_field(c, "dL_total_prev", "dL_total"));
        LPMB_TRY(copy_field(c, "TdL_total_prev", "TdL_total"));
        LPMB_TRY(run_geometry(c, v, "dLp0"));
        j2_energy_return_map_kernel<<<g, BT, 0, c->stream>>>(
            v, param(c, "particle_volume"), param(c, "J2_H"), param(c, "J2_xi"), param(c, "radius"), load_indicator, (const double *)ce->d,
            fptr<int>(c, "type"), fptr<double>(c, "sigmay"), Kn, broken, dL, dLt, fptr<double>(c, "dLp0"), fptr<double>(c, "J2_beta_eq0"),
            fptr<double>(c, "J2_alpha0"), fptr<double>(c, "dLp2"), fptr<double>(c, "J2_beta_eq2"), fptr<double>(c, "J2_alpha2"),
            fptr<double>(c, "ddLp"), fptr<double>(c, "J2_dlambda"), fptr<int>(c, "pl_flag"));
        LPMB_LAUNCH_CHECK(c);
        LPMB_TRY(run_geometry(c, v, "dLp2"));
        force_kernel<3><<<g, BT, 0, c->stream>>>(v, Kn, Tv, fptr<double>(c, "damage_D0"), dL, dLt, TdLt, csx, csy, csz, dL_ave, F, Pin, broken,
                                                 fptr<double>(c, "dL_total_prev"), fptr<double>(c, "TdL_total_prev"));
        LPMB_LAUNCH_CHECK(c);
    } else if (plmode == 5) {
        // J2, nonlinear isotropic hardening, in-place serial semantics (constitutive.c:689-863)
        LPMB_REQUIRE(c->params.count("J2_C"), LPMB_ERR_STATE, "J2_C not set");
        LPMB_REQUIRE(c->nn <= ISO_MAXNN, LPMB_ERR_UNSUPPORTED, "plmode 5: more than %d neighbours per particle", ISO_MAXNN);
        LPMB_REQUIRE(c->dim == 3, LPMB_ERR_UNSUPPORTED, "plmode 5 is a 3-D law (constitutive.c:706-708)");
        Field *ce = lpmb_field(c, "Ce");
        LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
        LPMB_TRY(iso_scratch(c));
        LPMB_TRY(copy_field(c, "dL_total_prev", "dL_total"));
        LPMB_TRY(copy_field(c, "TdL_total_prev", "TdL_total"));
        LPMB_TRY(copy_field(c, "dL_prev", "dL"));
        j2iso_trajectory_kernel<<<g, BT, 0, c->stream>>>(
            v, param(c, "particle_volume"), param(c, "J2_C"), (const double *)ce->d, fptr<int>(c, "type"), Kn, Tv, broken,
            fptr<double>(c, "distance_initial"), fptr<double>(c, "dLp0"), fptr<double>(c, "J2_alpha0"), fptr<double>(c, "J2_beta0"),
            fptr<double>(c, "J2_dlambda"), fptr<double>(c, "ddLp"), dL, dLt, TdLt, csx, csy, csz, F, fptr<double>(c, "iso_snap_dL"),
            fptr<double>(c, "iso_snap_t"), fptr<double>(c, "iso_self_dL"), fptr<double>(c, "iso_self_t"), fptr<int>(c, "iso_self_last"), -1);
        LPMB_LAUNCH_CHECK(c);
        j2iso_force_kernel<<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, broken, fptr<double>(c, "iso_snap_dL"), fptr<double>(c, "iso_snap_t"),
                                                     fptr<double>(c, "iso_self_dL"), fptr<double>(c, "iso_self_t"), fptr<int>(c, "iso_self_last"),
                                                     fptr<double>(c, "dL_prev"), fptr<double>(c, "dL_total_prev"), fptr<double>(c, "TdL_total_prev"),
                                                     csx, csy, csz, dL_ave, F, Pin, -1);
        LPMB_LAUNCH_CHECK(c);
    } else if (plmode == 1) {
        // crystal plasticity (constitutive.c:866-1396): geometry -> Miehe return map -> geometry -> averaged force
        LPMB_TRY(run_geometry(c, v, "dLp0"));
        LPMB_TRY(lpmb_cp_return_map(c));
        // the memo the serial loop leaves behind: every particle's increments of this pass exist (constitutive.c:114-117,957)
        fill_int_kernel<<<g, BT, 0, c->stream>>>(fptr<int>(c, "state_v"), c->N, 1);
        LPMB_LAUNCH_CHECK(c);
        LPMB_TRY(run_geometry(c, v, "dLp2"));
        force_kernel<0><<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, dL, dLt, TdLt, csx, csy, csz, dL_ave, F, Pin);
        LPMB_LAUNCH_CHECK(c);
    } else {
        lpmb_set_error("computeBondForceGeneral: plmode %d is not a law of the reference (0, 1, 3, 4, 5, 6 are)", plmode);
        return LPMB_ERR_UNSUPPORTED;
    }
    LPMB_TRY(lpmb_compute_stress(c));
    return lpmb_switch_state(c, 2);
}

// ---------------------------------------------------------------------------------------------
// per-particle entry points   computeBondForceElastic(ii)             constitutive.c:228-283
//                             computeBondForceIncrementalUpdating(ii) constitutive.c:167-225
//                             computeBondForceJ2mixedLinear3D(ii)     constitutive.c:466-686
//                             computeBondForceJ2energyReturnMap(ii,t) constitutive.c:286-463
//                             computeBondForceJ2nonlinearIso(ii)      constitutive.c:689-863
// ---------------------------------------------------------------------------------------------
// The reference evaluates particle ii together with its "star" (ii + the neighbours across intact bonds): the
// geometry / return-map passes rewrite dL, dL_total, TdL_total, cs* (ddLp, pl_flag) of EVERY star member, the force
// pass writes F, Pin (and the slot-[2] state, J2_dlambda, dL_ave, stress_tensor := 0) of ii only; nothing else may
// change.  Every pass is a pure per-particle function of xyz and the slot-[0] state, so the whole-lattice kernels
// above produce, for the star's rows, exactly the values of the restricted loops: they are run into scratch twins
// of the output fields (initialised with the current contents) and only the star's / ii's rows are committed.
// O(N) per call -- this exists for API completeness (nothing in the reference calls these once stiffness.c is
// replaced by the assembly kernel), not for speed.
// plmode 3: same scheme, but the star's geometry rows are committed BEFORE the force pass of ii, which then reads the
// live fields: across a broken bond the partner is not a star member and the reference reads its rows as they are.
// plmode 5: the law advances slot [0] of the whole star in place; the trajectory kernel restricted to the single
// caller ii does exactly that on the live fields (threads outside the star return untouched), the force kernel
// restricted to row ii reads the star's snapshots for caller ii and, across broken bonds, the live rows.
__global__ void fill_int_kernel(int *__restrict__ p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}

// xdLp = dLp[.][.][0] + ddLp over the initial bond list (constitutive.c:950-951 / 1308-1309)
__global__ void __launch_bounds__(BT)
cp_xdlp_kernel(BondView v, const double *__restrict__ dLp0, const double *__restrict__ ddLp, double *__restrict__ out)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const int n = v.nbi[i];
    for (int j = 0; j < v.nn; j++) {
        const size_t e = (size_t)j * v.Np + i;
        double x = dLp0[e];
        if (j < n)
            x += ddLp[e];
        out[e] = x;
    }
}

// slot [2] of particle ii when its increments are REUSED from the memo (constitutive.c:946-956, 1371-1379)
__global__ void cp_slot2_reuse_kernel(int ii, int Np, int nn, int S, const double *__restrict__ broken, const double *__restrict__ xdLp,
                                      const double *__restrict__ gy0, const double *__restrict__ dgy, const double *__restrict__ As0,
                                      const double *__restrict__ dAs, const double *__restrict__ A0, const double *__restrict__ dA,
                                      double *__restrict__ dLp2, double *__restrict__ gy2, double *__restrict__ As2, double *__restrict__ A2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nn) {
        const size_t e = (size_t)t * Np + ii;
        dLp2[e] = broken[e] * xdLp[e];
    }
    if (t < S) {
        const size_t e = (size_t)t * Np + ii;
        gy2[e] = gy0[e] + dgy[e];
        As2[e] = As0[e] + dAs[e];
    }
    if (t == 0)
        A2[ii] = A0[ii] + dA[ii];
}

template <typename T>
__global__ void commit_rows_kernel(T *__restrict__ dst, const T *__restrict__ src, const int *__restrict__ rows, int nrows, int comps, int Np)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrows * comps)
        return;
    const size_t e = (size_t)(t / nrows) * Np + rows[t % nrows];
    dst[e] = src[e];
}

static int pp_twin(lpmb_ctx *c, const char *name, void **out)
{
    Field *f = lpmb_field(c, name);
    LPMB_REQUIRE(f, LPMB_ERR_STATE, "field %s missing", name);
    const std::string tn = std::string("pp.") + name;
    if (!c->fields.count(tn))
        LPMB_TRY(lpmb_field_alloc(c, tn.c_str(), f->kind, f->type, f->comps));
    f = lpmb_field(c, name);
    Field *t = &c->fields[tn];
    LPMB_CUDA(cudaMemcpyAsync(t->d, f->d, f->count * f->elem(), cudaMemcpyDeviceToDevice, c->stream));
    *out = t->d;
    return LPMB_OK;
}

static int pp_commit(lpmb_ctx *c, const char *name, const int *d_rows, int nrows)
{
    Field *f = lpmb_field(c, name);
    Field *t = &c->fields[std::string("pp.") + name];
    const int work = nrows * f->comps;
    if (f->type == FT_F64)
        commit_rows_kernel<double><<<lpmb_blocks(work, 128), 128, 0, c->stream>>>((double *)f->d, (const double *)t->d, d_rows, nrows, f->comps, c->Np);
    else if (f->type == FT_I32)
        commit_rows_kernel<int><<<lpmb_blocks(work, 128), 128, 0, c->stream>>>((int *)f->d, (const int *)t->d, d_rows, nrows, f->comps, c->Np);
    else
        commit_rows_kernel<signed char><<<lpmb_blocks(work, 128), 128, 0, c->stream>>>((signed char *)f->d, (const signed char *)t->d, d_rows,
                                                                                     nrows, f->comps, c->Np);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

extern "C" int lpmb_bond_force_particle(lpmb_ctx *c, int plmode, int ii, int load_indicator)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_REQUIRE(ii >= 0 && ii < c->N, LPMB_ERR_ARG, "particle index %d out of range", ii);
    LPMB_REQUIRE(c->world == 1, LPMB_ERR_UNSUPPORTED, "per-particle laws are single-GPU only");
    LPMB_REQUIRE(plmode == 6 || plmode == 4 || plmode == 0 || plmode == 1 || plmode == 3 || plmode == 5, LPMB_ERR_UNSUPPORTED,
                 "per-particle evaluation: plmode %d is not a law of the reference (0, 1, 3, 4, 5, 6 are)", plmode);
    BondView v;
    LPMB_TRY(make_view(c, v));
    const int Np = c->Np, nn = c->nn;
    double *broken = fptr<double>(c, "damage_broken");
    LPMB_REQUIRE(broken, LPMB_ERR_STATE, "damage_broken missing");
    // the star: ii, then the neighbours across intact bonds in slot order, nb[ii] + 1 entries (constitutive.c:231-237)
    std::vector<int> hn(nn);
    std::vector<double> hb(nn);
    int nb_ii = 0;
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    LPMB_CUDA(cudaMemcpy2DAsync(hn.data(), sizeof(int), v.nbr + ii, (size_t)Np * sizeof(int), sizeof(int), nn, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaMemcpy2DAsync(hb.data(), sizeof(double), broken + ii, (size_t)Np * sizeof(double), sizeof(double), nn, cudaMemcpyDeviceToHost, c->stream));
    LPMB_D2H(c, &nb_ii, v.nb + ii, sizeof(int));
    std::vector<int> star(1, ii);
    for (int k = 0; k < nn; k++)
        if (hb[k] > LPMB_EPS && hn[k] != -1)
            star.push_back(hn[k]);
    LPMB_REQUIRE((int)star.size() == nb_ii + 1, LPMB_ERR_STATE, "particle %d: nb = %d but %d intact bonds (the reference would overrun its list)", ii,
                 nb_ii, (int)star.size() - 1);
    int *d_rows = nullptr;
    LPMB_CUDA(cudaMalloc(&d_rows, star.size() * sizeof(int)));
    LPMB_CUDA(cudaMemcpyAsync(d_rows, star.data(), star.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const int ns = (int)star.size();
    const int g = lpmb_blocks(c->N, BT);
    double *Kn = fptr<double>(c, "Kn"), *Tv = fptr<double>(c, "Tv"), *w = fptr<double>(c, "damage_w"), *L0 = fptr<double>(c, "distance_initial");
    int rc = LPMB_OK;
    auto run = [&]() -> int {
        if (plmode == 5) {
            LPMB_REQUIRE(c->params.count("J2_C") && c->params.count("particle_volume"), LPMB_ERR_STATE, "J2_C / particle_volume not set");
            LPMB_REQUIRE(c->nn <= ISO_MAXNN, LPMB_ERR_UNSUPPORTED, "plmode 5: more than %d neighbours per particle", ISO_MAXNN);
            LPMB_REQUIRE(c->dim == 3, LPMB_ERR_UNSUPPORTED, "plmode 5 is a 3-D law (constitutive.c:706-708)");
            Field *ce = lpmb_field(c, "Ce");
            LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
            LPMB_TRY(iso_scratch(c));
            double *dL = fptr<double>(c, "dL"), *dLt = fptr<double>(c, "dL_total"), *TdLt = fptr<double>(c, "TdL_total");
            double *csx = fptr<double>(c, "csx"), *csy = fptr<double>(c, "csy"), *csz = fptr<double>(c, "csz");
            j2iso_trajectory_kernel<<<g, BT, 0, c->stream>>>(
                v, param(c, "particle_volume"), param(c, "J2_C"), (const double *)ce->d, fptr<int>(c, "type"), Kn, Tv, broken, L0,
                fptr<double>(c, "dLp0"), fptr<double>(c, "J2_alpha0"), fptr<double>(c, "J2_beta0"), fptr<double>(c, "J2_dlambda"),
                fptr<double>(c, "ddLp"), dL, dLt, TdLt, csx, csy, csz, fptr<double>(c, "F"), fptr<double>(c, "iso_snap_dL"),
                fptr<double>(c, "iso_snap_t"), fptr<double>(c, "iso_self_dL"), fptr<double>(c, "iso_self_t"), fptr<int>(c, "iso_self_last"), ii);
            LPMB_LAUNCH_CHECK(c);
            // memset(stress_tensor[ii], 0, ...)  constitutive.c:834
            double *st = fptr<double>(c, "stress_tensor");
            LPMB_REQUIRE(st, LPMB_ERR_STATE, "stress_tensor missing");
            LPMB_CUDA(cudaMemset2DAsync(st + ii, (size_t)Np * sizeof(double), 0, sizeof(double), 6, c->stream));
            j2iso_force_kernel<<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, broken, fptr<double>(c, "iso_snap_dL"), fptr<double>(c, "iso_snap_t"),
                                                         fptr<double>(c, "iso_self_dL"), fptr<double>(c, "iso_self_t"),
                                                         fptr<int>(c, "iso_self_last"), dL, dLt, TdLt, csx, csy, csz, fptr<double>(c, "dL_ave"),
                                                         fptr<double>(c, "F"), fptr<double>(c, "Pin"), ii);
            LPMB_LAUNCH_CHECK(c);
            return LPMB_OK;
        }
        void *tF, *tPin;
        LPMB_TRY(pp_twin(c, "F", &tF));
        LPMB_TRY(pp_twin(c, "Pin", &tPin));
        if (plmode == 1) {
            // computeBondForceCPMiehe(ii), constitutive.c:866-1396.  Star members whose memo flag state_v is still 0 are
            // return-mapped (increments, cp_RSS, cp_Jact, pl_flag written, flag set); the others REUSE the increments an
            // earlier call left (:946-956).  Then the star's geometry with dLp[0] + ddLp, the force pass of ii over the live
            // fields, slot [2] of ii.
            const int S = (int)param(c, "nslipSys", 0.0);
            LPMB_REQUIRE(S > 0, LPMB_ERR_STATE, "plmode 1 needs the slip systems (lpmb_set_schmid_tensor)");
            int *state_v = fptr<int>(c, "state_v");
            LPMB_REQUIRE(state_v, LPMB_ERR_STATE, "state_v missing");
            std::vector<int> hsv(c->N), rows0;
            LPMB_CUDA(cudaMemcpyAsync(hsv.data(), state_v, (size_t)c->N * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            for (int q : star)
                if (hsv[q] == 0) {
                    rows0.push_back(q);
                    hsv[q] = 1;  // a particle listed twice would be reused the second time, as in the reference
                }
            const bool ii_fresh = !rows0.empty() && rows0[0] == ii;
            void *tdL, *tcx, *tcy, *tcz, *tdLt, *tTdLt, *tave;
            LPMB_TRY(pp_twin(c, "dL", &tdL));
            LPMB_TRY(pp_twin(c, "csx", &tcx));
            LPMB_TRY(pp_twin(c, "csy", &tcy));
            LPMB_TRY(pp_twin(c, "csz", &tcz));
            LPMB_TRY(pp_twin(c, "dL_total", &tdLt));
            LPMB_TRY(pp_twin(c, "TdL_total", &tTdLt));
            LPMB_TRY(pp_twin(c, "dL_ave", &tave));
            geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, fptr<double>(c, "dLp0"), broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy,
                                                        (double *)tcz, (double *)tdLt, (double *)tTdLt, nullptr);
            LPMB_LAUNCH_CHECK(c);
            if (!rows0.empty()) {
                CPIO io;
                void *q[11];
                const char *names[11] = {"dLp2", "cp_gy2", "cp_A2", "cp_A_single2", "ddLp", "cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single",
                                         "pl_flag"};
                for (int k = 0; k < 11; k++)
                    LPMB_TRY(pp_twin(c, names[k], &q[k]));
                io.dL = (double *)tdL, io.dLt = (double *)tdLt, io.TdLt = (double *)tTdLt;
                io.csx = (double *)tcx, io.csy = (double *)tcy, io.csz = (double *)tcz;
                io.dLp2 = (double *)q[0], io.gy2 = (double *)q[1], io.A2 = (double *)q[2], io.As2 = (double *)q[3], io.ddLp = (double *)q[4];
                io.RSS = (double *)q[5], io.Jact = (int *)q[6], io.dgy = (double *)q[7], io.dA = (double *)q[8], io.dAs = (double *)q[9];
                io.pl_flag = (int *)q[10];
                int err = 0;
                LPMB_TRY(lpmb_cp_return_map_io(c, &io, &err));
                for (int r0 : rows0)
                    LPMB_REQUIRE(err - 1 != r0, LPMB_ERR_STATE,
                                 "crystal plasticity: singular slip Jacobian at particle %d (the reference exits here, constitutive.c:1216-1221)", r0);
                int *d_rows0 = nullptr;
                LPMB_CUDA(cudaMalloc(&d_rows0, rows0.size() * sizeof(int)));
                LPMB_CUDA(cudaMemcpyAsync(d_rows0, rows0.data(), rows0.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                int rc2 = LPMB_OK;
                for (const char *n : {"ddLp", "cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single", "pl_flag"})
                    if (rc2 == LPMB_OK)
                        rc2 = pp_commit(c, n, d_rows0, (int)rows0.size());
                cudaStreamSynchronize(c->stream);
                cudaFree(d_rows0);
                LPMB_TRY(rc2);
                LPMB_CUDA(cudaMemcpyAsync(state_v, hsv.data(), (size_t)c->N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                LPMB_CUDA(cudaStreamSynchronize(c->stream));  // hsv is pageable host memory
            }
            // plastic stretch the geometry pass uses: previous + the (new or reused) increment, live arrays
            if (!c->fields.count("pp.xdLp"))
                LPMB_TRY(lpmb_field_alloc(c, "pp.xdLp", FK_BOND, FT_F64, c->nn));
            double *xdLp = fptr<double>(c, "pp.xdLp");
            cp_xdlp_kernel<<<g, BT, 0, c->stream>>>(v, fptr<double>(c, "dLp0"), fptr<double>(c, "ddLp"), xdLp);
            LPMB_LAUNCH_CHECK(c);
            geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, xdLp, broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy, (double *)tcz,
                                                        (double *)tdLt, (double *)tTdLt, nullptr);
            LPMB_LAUNCH_CHECK(c);
            for (const char *n : {"dL", "csx", "csy", "csz", "dL_total", "TdL_total"})
                LPMB_TRY(pp_commit(c, n, d_rows, ns));
            force_kernel<0><<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, fptr<double>(c, "dL"), fptr<double>(c, "dL_total"), fptr<double>(c, "TdL_total"),
                                                     fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"), (double *)tave,
                                                     (double *)tF, (double *)tPin);
            LPMB_LAUNCH_CHECK(c);
            LPMB_TRY(pp_commit(c, "dL_ave", d_rows, 1));
            if (ii_fresh) {
                for (const char *n : {"dLp2", "cp_gy2", "cp_A2", "cp_A_single2"})
                    LPMB_TRY(pp_commit(c, n, d_rows, 1));
            } else {
                const int T = c->nn > S ? c->nn : S;
                cp_slot2_reuse_kernel<<<lpmb_blocks(T, 64), 64, 0, c->stream>>>(
                    ii, Np, c->nn, S, broken, xdLp, fptr<double>(c, "cp_gy0"), fptr<double>(c, "cp_dgy"), fptr<double>(c, "cp_A_single0"),
                    fptr<double>(c, "cp_dA_single"), fptr<double>(c, "cp_A0"), fptr<double>(c, "cp_dA"), fptr<double>(c, "dLp2"),
                    fptr<double>(c, "cp_gy2"), fptr<double>(c, "cp_A_single2"), fptr<double>(c, "cp_A2"));
                LPMB_LAUNCH_CHECK(c);
            }
        } else if (plmode == 3) {
            LPMB_REQUIRE(c->params.count("J2_H") && c->params.count("J2_xi") && c->params.count("radius") && c->params.count("particle_volume"),
                         LPMB_ERR_STATE, "J2_H / J2_xi / radius / particle_volume not set");
            Field *ce = lpmb_field(c, "Ce");
            LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
            void *tdL, *tcx, *tcy, *tcz, *tdLt, *tTdLt, *tave, *tdLp2, *tb2, *ta2, *tddLp, *tdl, *tpf;
            LPMB_TRY(pp_twin(c, "dL", &tdL));
            LPMB_TRY(pp_twin(c, "csx", &tcx));
            LPMB_TRY(pp_twin(c, "csy", &tcy));
            LPMB_TRY(pp_twin(c, "csz", &tcz));
            LPMB_TRY(pp_twin(c, "dL_total", &tdLt));
            LPMB_TRY(pp_twin(c, "TdL_total", &tTdLt));
            LPMB_TRY(pp_twin(c, "dL_ave", &tave));
            LPMB_TRY(pp_twin(c, "dLp2", &tdLp2));
            LPMB_TRY(pp_twin(c, "J2_beta_eq2", &tb2));
            LPMB_TRY(pp_twin(c, "J2_alpha2", &ta2));
            LPMB_TRY(pp_twin(c, "ddLp", &tddLp));
            LPMB_TRY(pp_twin(c, "J2_dlambda", &tdl));
            LPMB_TRY(pp_twin(c, "pl_flag", &tpf));
            geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, fptr<double>(c, "dLp0"), broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy,
                                                        (double *)tcz, (double *)tdLt, (double *)tTdLt, nullptr);
            LPMB_LAUNCH_CHECK(c);
            j2_energy_return_map_kernel<<<g, BT, 0, c->stream>>>(
                v, param(c, "particle_volume"), param(c, "J2_H"), param(c, "J2_xi"), param(c, "radius"), load_indicator, (const double *)ce->d,
                fptr<int>(c, "type"), fptr<double>(c, "sigmay"), Kn, broken, (double *)tdL, (double *)tdLt, fptr<double>(c, "dLp0"),
                fptr<double>(c, "J2_beta_eq0"), fptr<double>(c, "J2_alpha0"), (double *)tdLp2, (double *)tb2, (double *)ta2, (double *)tddLp,
                (double *)tdl, (int *)tpf);
            LPMB_LAUNCH_CHECK(c);
            geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, (double *)tdLp2, broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy,
                                                        (double *)tcz, (double *)tdLt, (double *)tTdLt, nullptr);
            LPMB_LAUNCH_CHECK(c);
            // the star's rows become visible first ...
            for (const char *n : {"dL", "csx", "csy", "csz", "dL_total", "TdL_total", "ddLp", "pl_flag"})
                LPMB_TRY(pp_commit(c, n, d_rows, ns));
            // ... then the force pass of ii over the LIVE fields (`prev` = the same arrays: either branch reads memory as it is)
            double *dL = fptr<double>(c, "dL"), *dLt = fptr<double>(c, "dL_total"), *TdLt = fptr<double>(c, "TdL_total");
            force_kernel<3><<<g, BT, 0, c->stream>>>(v, Kn, Tv, fptr<double>(c, "damage_D0"), dL, dLt, TdLt, fptr<double>(c, "csx"),
                                                     fptr<double>(c, "csy"), fptr<double>(c, "csz"), (double *)tave, (double *)tF, (double *)tPin,
                                                     broken, dLt, TdLt);
            LPMB_LAUNCH_CHECK(c);
            for (const char *n : {"dL_ave", "dLp2", "J2_beta_eq2", "J2_alpha2", "J2_dlambda"})
                LPMB_TRY(pp_commit(c, n, d_rows, 1));
        } else if (plmode == 4) {
            void *tddL, *tddLt, *tTddLt;
            LPMB_TRY(pp_twin(c, "ddL", &tddL));
            LPMB_TRY(pp_twin(c, "ddL_total", &tddLt));
            LPMB_TRY(pp_twin(c, "TddL_total", &tTddLt));
            predictor_geometry_kernel<<<g, BT, 0, c->stream>>>(v, fptr<double>(c, "xyz_temp"), broken, Tv, (double *)tddL, (double *)tddLt,
                                                               (double *)tTddLt);
            LPMB_LAUNCH_CHECK(c);
            predictor_force_kernel<<<g, BT, 0, c->stream>>>(v, Kn, Tv, broken, fptr<double>(c, "F_temp"), (double *)tddL, (double *)tddLt,
                                                            (double *)tTddLt, fptr<double>(c, "csx"), fptr<double>(c, "csy"), fptr<double>(c, "csz"),
                                                            (double *)tF, (double *)tPin);
            LPMB_LAUNCH_CHECK(c);
            for (const char *n : {"ddL", "ddL_total", "TddL_total"})
                LPMB_TRY(pp_commit(c, n, d_rows, ns));
        } else {
            void *tdL, *tcx, *tcy, *tcz, *tdLt, *tTdLt, *tave;
            LPMB_TRY(pp_twin(c, "dL", &tdL));
            LPMB_TRY(pp_twin(c, "csx", &tcx));
            LPMB_TRY(pp_twin(c, "csy", &tcy));
            LPMB_TRY(pp_twin(c, "csz", &tcz));
            LPMB_TRY(pp_twin(c, "dL_total", &tdLt));
            LPMB_TRY(pp_twin(c, "TdL_total", &tTdLt));
            LPMB_TRY(pp_twin(c, "dL_ave", &tave));
            geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, fptr<double>(c, "dLp0"), broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy,
                                                        (double *)tcz, (double *)tdLt, (double *)tTdLt, nullptr);
            LPMB_LAUNCH_CHECK(c);
            if (plmode == 6) {
                force_kernel<6><<<g, BT, 0, c->stream>>>(v, Kn, Tv, broken, (double *)tdL, (double *)tdLt, (double *)tTdLt, (double *)tcx,
                                                         (double *)tcy, (double *)tcz, (double *)tave, (double *)tF, (double *)tPin);
                LPMB_LAUNCH_CHECK(c);
            } else {
                LPMB_REQUIRE(c->params.count("J2_H") && c->params.count("J2_xi") && c->params.count("particle_volume"), LPMB_ERR_STATE,
                             "J2_H / J2_xi / particle_volume not set");
                Field *ce = lpmb_field(c, "Ce");
                LPMB_REQUIRE(ce, LPMB_ERR_STATE, "Ce not uploaded (lpmb_calc_kntv)");
                void *tdLp2, *tb2, *ta2, *tddLp, *tdl, *tpf;
                LPMB_TRY(pp_twin(c, "dLp2", &tdLp2));
                LPMB_TRY(pp_twin(c, "J2_beta2", &tb2));
                LPMB_TRY(pp_twin(c, "J2_alpha2", &ta2));
                LPMB_TRY(pp_twin(c, "ddLp", &tddLp));
                LPMB_TRY(pp_twin(c, "J2_dlambda", &tdl));
                LPMB_TRY(pp_twin(c, "pl_flag", &tpf));
                j2_return_map_kernel<<<g, BT, 0, c->stream>>>(
                    v, param(c, "particle_volume"), param(c, "J2_H"), param(c, "J2_xi"), (const double *)ce->d, fptr<int>(c, "type"),
                    fptr<double>(c, "sigmay"), Kn, Tv, w, broken, L0, (double *)tdL, (double *)tdLt, (double *)tTdLt, (double *)tcx, (double *)tcy,
                    (double *)tcz, fptr<double>(c, "dLp0"), fptr<double>(c, "J2_beta0"), fptr<double>(c, "J2_alpha0"), (double *)tdLp2, (double *)tb2,
                    (double *)ta2, (double *)tddLp, (double *)tdl, (int *)tpf);
                LPMB_LAUNCH_CHECK(c);
                geometry_kernel<0><<<g, BT, 0, c->stream>>>(v, L0, (double *)tdLp2, broken, Tv, (double *)tdL, (double *)tcx, (double *)tcy,
                                                            (double *)tcz, (double *)tdLt, (double *)tTdLt, nullptr);
                LPMB_LAUNCH_CHECK(c);
                force_kernel<0><<<g, BT, 0, c->stream>>>(v, Kn, Tv, w, (double *)tdL, (double *)tdLt, (double *)tTdLt, (double *)tcx, (double *)tcy,
                                                         (double *)tcz, (double *)tave, (double *)tF, (double *)tPin);
                LPMB_LAUNCH_CHECK(c);
                for (const char *n : {"ddLp", "pl_flag"})
                    LPMB_TRY(pp_commit(c, n, d_rows, ns));
                for (const char *n : {"dL_ave", "dLp2", "J2_beta2", "J2_alpha2", "J2_dlambda"})
                    LPMB_TRY(pp_commit(c, n, d_rows, 1));
                // memset(stress_tensor[ii], 0, ...)  constitutive.c:647
                double *st = fptr<double>(c, "stress_tensor");
                LPMB_REQUIRE(st, LPMB_ERR_STATE, "stress_tensor missing");
                LPMB_CUDA(cudaMemset2DAsync(st + ii, (size_t)Np * sizeof(double), 0, sizeof(double), 6, c->stream));
            }
            for (const char *n : {"dL", "csx", "csy", "csz", "dL_total", "TdL_total"})
                LPMB_TRY(pp_commit(c, n, d_rows, ns));
        }
        LPMB_TRY(pp_commit(c, "F", d_rows, 1));
        LPMB_TRY(pp_commit(c, "Pin", d_rows, 1));
        return LPMB_OK;
    };
    rc = run();
    cudaStreamSynchronize(c->stream);
    cudaFree(d_rows);
    return rc;
}

extern "C" int lpmb_update_crack(lpmb_ctx *c)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    BondView v;
    LPMB_TRY(make_view(c, v));
    update_crack_kernel<<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(
        v, fptr<double>(c, "damage_broken"), fptr<double>(c, "damage_w"), fptr<double>(c, "csx"), fptr<double>(c, "csy"),
        fptr<double>(c, "csz"), fptr<double>(c, "F"), fptr<double>(c, "Pin"), fptr<double>(c, "damage_visual"), fptr<int>(c, "fix_index"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

extern "C" int lpmb_update_rr(lpmb_ctx *c, double *norm_residual, double *norm_reaction)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_TRY(lpmb_cg_alloc(c));
    const int grid = c->sm_count * 4;
    LPMB_REQUIRE(2 * grid + 2 <= 2 * c->cg.max_blocks, LPMB_ERR_STATE, "partials buffer too small");
    double *partials = c->cg.partials;
    update_rr_kernel<<<grid, 256, 0, c->stream>>>(c->dim, c->Np, c->N, lpmb_own0(c), lpmb_own1(c), fptr<int>(c, "dispBC_index"),
                                                    fptr<double>(c, "Pex"), fptr<double>(c, "Pin"), fptr<double>(c, "residual"), partials);
    LPMB_LAUNCH_CHECK(c);
    double *out2 = c->cg.scal + 12;
    finish_rr_kernel<<<1, 256, 0, c->stream>>>(partials, grid, out2);
    LPMB_LAUNCH_CHECK(c);
    LPMB_TRY(lpmb_dist_allreduce_sum(c, out2, 2));
    if (norm_residual || norm_reaction) {
        LPMB_CUDA(cudaMemcpyAsync(c->cg.h_scal + 12, out2, 16, cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        if (norm_residual)
            *norm_residual = sqrt(c->cg.h_scal[12]);
        if (norm_reaction)
            *norm_reaction = sqrt(c->cg.h_scal[13]);
    }
    return LPMB_OK;
}

// calcKnTv   stiffness.c:11-268: KnTve[type] = (radius) * M_lattice * Ce[type]; per bond by shell
__global__ void __launch_bounds__(BT)
kntv_kernel(BondView v, int lattice, const double *__restrict__ KnTve /* [ntype][3] */, const int *__restrict__ type,
            double *__restrict__ Kn, double *__restrict__ Tv)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= v.N)
        return;
    const size_t Np = v.Np;
    const int n = v.nbi[i], ti = type[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int s = v.nsign[e];
        if (lattice == LPMB_LATTICE_SC) {  // stiffness.c:192-201: average of the two end particles' types
            const int tj = type[v.nbr[e]];
            Kn[e] = 0.5 * (KnTve[3 * ti + s] + KnTve[3 * tj + s]);
            Tv[e] = 0.5 * (KnTve[3 * ti + 2] + KnTve[3 * tj + 2]);
        } else if (lattice == LPMB_LATTICE_HEX) {  // stiffness.c:58-65 (2 columns)
            Kn[e] = KnTve[3 * ti + 0];
            Tv[e] = KnTve[3 * ti + 1];
        } else {  // square, FCC, BCC: stiffness.c:25-41, 218-234, 250-266
            Kn[e] = KnTve[3 * ti + s];
            Tv[e] = KnTve[3 * ti + 2];
        }
    }
}

extern "C" int lpmb_calc_kntv(lpmb_ctx *c, const double *Ce, int ntype)
{
    LPMB_REQUIRE(c && Ce && ntype > 0, LPMB_ERR_ARG, "lpmb_calc_kntv: bad argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_REQUIRE(c->params.count("radius"), LPMB_ERR_STATE, "parameter radius not set");
    const double radius = param(c, "radius");
    // the 3x3 (2x2) constant maps of stiffness.c:17-20, 51-53, 179-182, 210-213, 242-245; the product is
    // evaluated like the reference's row-major dgemm: alpha * (sum_p M[r][p] * Ce[p])
    std::vector<double> knt((size_t)ntype * 3, 0.0);
    for (int k = 0; k < ntype; k++) {
        const double *ce = Ce + 3 * k;
        double M[9];
        double alpha = radius;
        int rows = 3;
        switch (c->lattice) {
        case LPMB_LATTICE_SQUARE: {
            const double m[9] = {1 / 2.0, -1 / 2.0, 0.0, 0.0, 0.0, 1 / 2.0, 0.0, 1.0 / 12.0, -1.0 / 12.0};
            memcpy(M, m, sizeof(m));
            alpha = 1.0;
            break;
        }
        case LPMB_LATTICE_HEX: {
            const double m[9] = {sqrt(3.0) / 12.0, -sqrt(3.0) / 12.0, 0, -sqrt(3.0) / 144.0, sqrt(3.0) / 48.0, 0, 0, 0, 0};
            memcpy(M, m, sizeof(m));
            alpha = 1.0;
            rows = 2;
            break;
        }
        case LPMB_LATTICE_SC: {
            const double m[9] = {1, -1, -1, 0, 0, 1, 0, 1.0 / 18.0, -1.0 / 18.0};
            memcpy(M, m, sizeof(m));
            break;
        }
        case LPMB_LATTICE_FCC: {
            const double m[9] = {0, 0, sqrt(2.0), sqrt(2.0) / 4.0, -sqrt(2.0) / 4.0, -sqrt(2.0) / 4.0, 0, sqrt(2.0) / 24.0, -sqrt(2.0) / 24.0};
            memcpy(M, m, sizeof(m));
            break;
        }
        case LPMB_LATTICE_BCC: {
            const double m[9] = {0., 0., sqrt(3.0), 1. / sqrt(3.0), -1. / sqrt(3.0), 0., 0., sqrt(3.0) / 14.0, sqrt(3.0) / 14.0};
            memcpy(M, m, sizeof(m));
            break;
        }
        default:
            lpmb_set_error("calcKnTv: lattice %d", c->lattice);
            return LPMB_ERR_ARG;
        }
        for (int r = 0; r < rows; r++) {
            double s = 0.0;
            if (rows == 2) {
                for (int p = 0; p < 2; p++)
                    s += M[r * 3 + p] * ce[p];
            } else {
                for (int p = 0; p < 3; p++)
                    s += M[r * 3 + p] * ce[p];
            }
            knt[3 * k + r] = alpha * s;
        }
    }
    // Ce and KnTve live as small raw device fields
    for (const char *nm : {"Ce", "KnTve"}) {
        auto it = c->fields.find(nm);
        if (it != c->fields.end() && it->second.count != (size_t)ntype * 3) {
            cudaFree(it->second.d);
            c->fields.erase(it);
        }
        if (!c->fields.count(nm)) {
            Field f;
            f.kind = FK_RAW;
            f.type = FT_F64;
            f.comps = 3;
            f.count = (size_t)ntype * 3;
            LPMB_CUDA(cudaMalloc(&f.d, f.count * 8));
            c->fields[nm] = f;
        }
    }
    LPMB_CUDA(cudaMemcpyAsync(c->fields["Ce"].d, Ce, (size_t)ntype * 24, cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaMemcpyAsync(c->fields["KnTve"].d, knt.data(), (size_t)ntype * 24, cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    BondView v;
    LPMB_TRY(make_view(c, v));
    kntv_kernel<<<lpmb_blocks(c->N, BT), BT, 0, c->stream>>>(v, c->lattice, (const double *)c->fields["KnTve"].d, fptr<int>(c, "type"),
                                                           fptr<double>(c, "Kn"), fptr<double>(c, "Tv"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}
