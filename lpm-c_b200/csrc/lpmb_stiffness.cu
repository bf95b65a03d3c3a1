// lpmb_stiffness.cu -- finite-difference elastic tangent, assembled straight into the device block
// matrix (compiled with -fmad=false: strict IEEE, bit-compatible with the reference's arithmetic).
//
// Replaces calcStiffness{2,3}DFiniteDifference(6), reference src/stiffness.c:271-516:
//   for every particle i: base computeBondForceElastic(i) (src/constitutive.c:228-283), then for every
//   conn particle c and coordinate r: xyz[c][r] += EPS*radius, re-evaluate, restore;
//   A_i[c][s][r] = (Pin_i^pert[s] - Pin_i^base[s]) / EPS / radius       (stiffness.c:419-427)
//   K_ii = upper triangle of A_i[i];  K_ij = 0.5*A_i[j] + 0.5*A_j[i]^T   (stiffness.c:441-481)
//
// Design.  One CTA per particle i, one thread per perturbation (c,r) (<= 183 in 3-D).  The bond star
// of i -- i, its <= nn neighbours and *their* bond lists with neighbour positions, L0, dLp, Tv,
// damage_broken -- is staged once in shared memory (~21 KB) together with the unperturbed bond
// stretches / direction cosines / dilatation sums.  A perturbation of c only changes the bonds that
// touch c, so each thread recomputes (with exactly the reference's expressions and summation order)
// only the star members whose bond list contains c -- found in O(1) from a 64-bit "affected" mask per
// star member -- and takes the cached unperturbed sums otherwise.  Unaffected quantities are pure
// functions of unchanged inputs, so the result equals the reference's brute-force re-evaluation of
// all 19x18 bonds bit for bit, at ~1/100 of its sqrt/div count.  To keep warps convergent the work is split in two
// phases: (1) every affected (neighbour a, conn q, coordinate r) triple is a task -- tasks of one neighbour are padded
// to a multiple of the warp size, so a warp walks one bond list with one skip pattern -- whose perturbed shell sums go
// to a compact shared-memory table; (2) one thread per perturbation assembles Pin from the table (O(1) lookup by the
// rank of q in the neighbour's affected mask); only the perturbations that move the owner's own bonds (conn members
// that are the owner or its neighbours: renumbered to the first two warps) recompute geometry there.
// The 3x3 (2x2) blocks A_i[c] go to
// row i of the SELL matrix; a second kernel symmetrises pairs in place (no atomics: every (i,j>i)
// pair is owned by one thread).
#include "lpmb_internal.cuh"

template <int D, int NN>
struct StarSmem {
    int sid[NN + 1];                 // star member particle ids (0 = owner), -1 = none
    int nbi[NN + 1];                 // nb_initial of each member
    unsigned long long amask[NN + 1];// bit q set <=> conn[i][q] is the member itself or one of its neighbours
    double pos[NN + 1][3];
    int nid[NN + 1][NN];
    double npos[NN + 1][NN][3];
    double L0[NN + 1][NN], dLp[NN + 1][NN], brk[NN + 1][NN], Tv[NN + 1][NN];
    signed char sg[NN + 1][NN];
    double bt[NN + 1][2], bT[NN + 1][2];  // unperturbed dL_total / TdL_total
    double bd[NN + 1][NN];                // unperturbed bond stretch of every member (bd[0] = the owner's)
    double bcs[NN][3];                    // unperturbed owner direction cosines
    unsigned char cq[NN + 1][NN];         // conn position of member a's neighbour mm (255: not a conn member of i)
    unsigned char selfq[NN + 1];          // conn position of the member itself
    double Kn[NN];
    int conn[64];
    int nbc;
    double base_pin[3];
    // phase-1 table: perturbed shell sums (shell of the owner's bond to a) of neighbour a for the rank-th set bit of amask[a]
    double ptj[NN * (NN + 1) * D], pTj[NN * (NN + 1) * D];
    int toff[NN + 2];
    unsigned char qord[64];               // perturbation order: conn members that move the owner's bonds first
};

template <int D, int NN, int T>
__global__ void __launch_bounds__(T)
fd_stiffness_kernel(int N, int Np, double h, double eps, double radius, const int *__restrict__ nbr, const signed char *__restrict__ nsign,
                    const int *__restrict__ nbi_g, const double *__restrict__ xyz, const double *__restrict__ L0g,
                    const double *__restrict__ dLp0g, const double *__restrict__ brkg, const double *__restrict__ Tvg,
                    const double *__restrict__ Kng, const long long *__restrict__ sptr, const int *__restrict__ col,
                    const int *__restrict__ nbc_g, double *__restrict__ val, double *__restrict__ F_side, double *__restrict__ Pin_side)
{
    extern __shared__ unsigned char smem_raw[];
    StarSmem<D, NN> &S = *reinterpret_cast<StarSmem<D, NN> *>(smem_raw);
    const int i = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t Npz = Np;
    const long long krow = sptr[i >> 5];
    const int lane_i = i & 31;

    // ---- stage the star ----
    if (tid == 0) {
        S.nbc = nbc_g[i];
        S.sid[0] = i;
    }
    if (tid < NN) {
        const int nj = (tid < nbi_g[i]) ? nbr[(size_t)tid * Npz + i] : -1;
        S.sid[tid + 1] = nj;
    }
    if (tid < 64)
        S.conn[tid] = (tid < nbc_g[i]) ? col[(krow + tid) * 32 + lane_i] : -1;
    if (tid < NN + 1) {
        S.amask[tid] = 0ull;
        S.selfq[tid] = 255;
    }
    __syncthreads();
    if (tid < NN + 1) {
        const int pid = S.sid[tid];
        S.nbi[tid] = pid >= 0 ? nbi_g[pid] : 0;
        if (pid >= 0) {
            S.pos[tid][0] = xyz[pid];
            S.pos[tid][1] = xyz[Npz + pid];
            S.pos[tid][2] = xyz[2 * Npz + pid];
        }
    }
    __syncthreads();
    const int nbc = S.nbc;
    for (int e = tid; e < (NN + 1) * NN; e += T) {
        const int a = e / NN, m = e % NN;
        const int pid = S.sid[a];
        int nj = -1;
        if (pid >= 0 && m < S.nbi[a]) {
            const size_t g = (size_t)m * Npz + pid;
            nj = nbr[g];
            S.npos[a][m][0] = xyz[nj];
            S.npos[a][m][1] = xyz[Npz + nj];
            S.npos[a][m][2] = xyz[2 * Npz + nj];
            S.L0[a][m] = L0g[g];
            S.dLp[a][m] = dLp0g[g];
            S.brk[a][m] = brkg[g];
            S.Tv[a][m] = Tvg[g];
            S.sg[a][m] = nsign[g];
            if (a == 0)
                S.Kn[m] = Kng[g];
        }
        S.nid[a][m] = nj;
        S.cq[a][m] = 255;
        // affected mask: position of nj (and of the member itself) in conn[i]
        if (nj >= 0) {
            int lo = 0, hi = nbc - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1, v = S.conn[mid];
                if (v == nj) {
                    atomicOr(&S.amask[a], 1ull << mid);
                    S.cq[a][m] = (unsigned char)mid;
                    break;
                }
                if (v < nj)
                    lo = mid + 1;
                else
                    hi = mid - 1;
            }
        }
        if (m == 0 && pid >= 0) {
            int lo = 0, hi = nbc - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1, v = S.conn[mid];
                if (v == pid) {
                    atomicOr(&S.amask[a], 1ull << mid);
                    S.selfq[a] = (unsigned char)mid;
                    break;
                }
                if (v < pid)
                    lo = mid + 1;
                else
                    hi = mid - 1;
            }
        }
    }
    __syncthreads();
    // ---- unperturbed geometry of every star member (constitutive.c:241-260): one thread per bond, then the
    //      shell sums by one thread per member in the reference's order ----
    for (int e = tid; e < (NN + 1) * NN; e += T) {
        const int a = e / NN, m = e % NN;
        if (S.sid[a] < 0 || m >= S.nbi[a])
            continue;
        const double dx = S.pos[a][0] - S.npos[a][m][0], dy = S.pos[a][1] - S.npos[a][m][1], dz = S.pos[a][2] - S.npos[a][m][2];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        double d = dis - S.L0[a][m];
        d -= S.dLp[a][m];
        d *= S.brk[a][m];
        S.bd[a][m] = d;
        if (a == 0) {
            S.bcs[m][0] = dx / dis;
            S.bcs[m][1] = dy / dis;
            S.bcs[m][2] = dz / dis;
        }
    }
    if (tid == 0) {
        int acc = 0;
        for (int a = 1; a <= S.nbi[0]; a++) {
            S.toff[a] = acc;
            acc += __popcll(S.amask[a]);
        }
    }
    if (tid < nbc) {
        // perturbation order: conn members in the owner's own mask (owner + its neighbours) first
        const unsigned long long own = S.amask[0], valid = nbc >= 64 ? ~0ull : ((1ull << nbc) - 1ull), bt = 1ull << tid;
        const int pos = (own & bt) ? __popcll(own & (bt - 1ull)) : __popcll(own & valid) + __popcll(~own & valid & (bt - 1ull));
        S.qord[pos] = (unsigned char)tid;
    }
    __syncthreads();
    if (tid < NN + 1 && S.sid[tid] >= 0) {
        const int a = tid;
        double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
        for (int m = 0; m < S.nbi[a]; m++) {
            const double d = S.bd[a][m];
            const double td = S.Tv[a][m] * d;
            if (S.sg[a][m] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
        }
        S.bt[a][0] = t0;
        S.bt[a][1] = t1;
        S.bT[a][0] = T0;
        S.bT[a][1] = T1;
    }
    __syncthreads();

    // ---- phase 1: perturbed shell sums of the neighbours (only the shell of the owner's bond to a is needed) ----
    // A perturbation of c changes, in neighbour a's list, either every bond (c == a: "self" tasks, one warp pair) or
    // exactly one bond (c is a neighbour of a: "one-bond" tasks); unchanged stretches are bit-identical to the
    // unperturbed ones, so a one-bond task evaluates one bond and re-adds the cached stretches in the reference's
    // order.  Task slots of one neighbour are padded to TPA, so a warp works on one neighbour.
    constexpr int TPA = (D == 3) ? 64 : 32;  // >= NN * D
    const int n0 = S.nbi[0];
    for (int ts = tid; ts < (n0 + 1) * TPA; ts += T) {
        if (ts >= n0 * TPA) {
            // self tasks (a, r), and -- on a thread that has nothing else to do -- the base internal force
            const int kk = ts - n0 * TPA;
            if (kk == TPA - 1) {
                // base internal force (constitutive.c:264-279)
                double q0 = 0.0, q1 = 0.0, q2 = 0.0;
                for (int m = 0; m < n0; m++) {
                    const int s = S.sg[0][m];
                    double f = 2.0 * S.Kn[m] * S.bd[0][m] + 0.5 * (S.bT[0][s] + S.bT[m + 1][s]) + 0.5 * S.Tv[0][m] * (S.bt[0][s] + S.bt[m + 1][s]);
                    f *= S.brk[0][m];
                    q0 += S.bcs[m][0] * f;
                    q1 += S.bcs[m][1] * f;
                    q2 += S.bcs[m][2] * f;
                }
                S.base_pin[0] = q0;
                S.base_pin[1] = q1;
                S.base_pin[2] = q2;
            }
            if (kk >= n0 * D)
                continue;
            const int a = kk / D + 1, r = kk % D;
            if (S.selfq[a] == 255)
                continue;  // (cannot happen for conn built by neighbor.c: every neighbour is a conn member)
            const int s = S.sg[0][a - 1];
            double apos[3] = {S.pos[a][0], S.pos[a][1], S.pos[a][2]};
            apos[r] = apos[r] + h;
            double t = 0, TT = 0;
            const int na = S.nbi[a];
            for (int mm = 0; mm < na; mm++) {
                if (S.sg[a][mm] != s)
                    continue;
                const double dx = apos[0] - S.npos[a][mm][0], dy = apos[1] - S.npos[a][mm][1], dz = apos[2] - S.npos[a][mm][2];
                const double dis = sqrt(dx * dx + dy * dy + dz * dz);
                double d = dis - S.L0[a][mm];
                d -= S.dLp[a][mm];
                d *= S.brk[a][mm];
                t += d;
                TT += S.Tv[a][mm] * d;
            }
            const unsigned long long bq = 1ull << S.selfq[a];
            const int idx = (S.toff[a] + __popcll(S.amask[a] & (bq - 1ull))) * D + r;
            S.ptj[idx] = t;
            S.pTj[idx] = TT;
            continue;
        }
        // one-bond tasks (a, mm, r)
        const int a = ts / TPA + 1, kk = ts % TPA;
        const int na = S.nbi[a];
        if (kk >= na * D)
            continue;
        const int mm = kk / D, r = kk % D;
        const int qc = S.cq[a][mm];
        if (qc == 255)
            continue;  // that neighbour of a is not a conn member of i: never perturbed
        const int s = S.sg[0][a - 1];
        double t, TT;
        if (S.sg[a][mm] != s) {
            t = S.bt[a][s];  // the moved bond is in the other shell: these sums do not change
            TT = S.bT[a][s];
        } else {
            double np[3] = {S.npos[a][mm][0], S.npos[a][mm][1], S.npos[a][mm][2]};
            np[r] = np[r] + h;
            const double dx = S.pos[a][0] - np[0], dy = S.pos[a][1] - np[1], dz = S.pos[a][2] - np[2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            double dn = dis - S.L0[a][mm];
            dn -= S.dLp[a][mm];
            dn *= S.brk[a][mm];
            t = 0;
            TT = 0;
            for (int m2 = 0; m2 < na; m2++) {
                if (S.sg[a][m2] != s)
                    continue;
                const double d = m2 == mm ? dn : S.bd[a][m2];
                t += d;
                TT += S.Tv[a][m2] * d;
            }
        }
        const unsigned long long bq = 1ull << qc;
        const int idx = (S.toff[a] + __popcll(S.amask[a] & (bq - 1ull))) * D + r;
        S.ptj[idx] = t;
        S.pTj[idx] = TT;
    }
    __syncthreads();

    // ---- phase 2: one perturbation per thread ----
    const int p = tid;
    if (p >= D * nbc)
        return;
    const int q = S.qord[p / D], r = p % D;
    const int c = S.conn[q];
    const unsigned long long bit = 1ull << q;

    // owner: the perturbation moves all of its bonds (c == i), one of them (c is a neighbour) or none
    double ti[2], Ti[2];
    double od[NN], ocx[NN], ocy[NN], ocz[NN];  // c == i only
    double dn = 0.0, ncx = 0.0, ncy = 0.0, ncz = 0.0;
    int mstar = -1;
    const bool own_aff = (S.amask[0] & bit) != 0ull;
    const bool own_all = c == i;
    if (own_all) {
        double opos[3] = {S.pos[0][0], S.pos[0][1], S.pos[0][2]};
        opos[r] = opos[r] + h;
        double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
        for (int m = 0; m < n0; m++) {
            const double dx = opos[0] - S.npos[0][m][0], dy = opos[1] - S.npos[0][m][1], dz = opos[2] - S.npos[0][m][2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            double d = dis - S.L0[0][m];
            d -= S.dLp[0][m];
            d *= S.brk[0][m];
            const double td = S.Tv[0][m] * d;
            if (S.sg[0][m] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
            od[m] = d;
            ocx[m] = dx / dis;
            ocy[m] = dy / dis;
            ocz[m] = dz / dis;
        }
        ti[0] = t0;
        ti[1] = t1;
        Ti[0] = T0;
        Ti[1] = T1;
    } else if (own_aff) {
        for (int m = 0; m < n0; m++)
            if (S.nid[0][m] == c)
                mstar = m;
        double np[3] = {S.npos[0][mstar][0], S.npos[0][mstar][1], S.npos[0][mstar][2]};
        np[r] = np[r] + h;
        const double dx = S.pos[0][0] - np[0], dy = S.pos[0][1] - np[1], dz = S.pos[0][2] - np[2];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        dn = dis - S.L0[0][mstar];
        dn -= S.dLp[0][mstar];
        dn *= S.brk[0][mstar];
        ncx = dx / dis;
        ncy = dy / dis;
        ncz = dz / dis;
        double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
        for (int m = 0; m < n0; m++) {
            const double d = m == mstar ? dn : S.bd[0][m];
            const double td = S.Tv[0][m] * d;
            if (S.sg[0][m] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
        }
        ti[0] = t0;
        ti[1] = t1;
        Ti[0] = T0;
        Ti[1] = T1;
    } else {
        ti[0] = S.bt[0][0];
        ti[1] = S.bt[0][1];
        Ti[0] = S.bT[0][0];
        Ti[1] = S.bT[0][1];
    }

    const bool last = (q == nbc - 1) && (r == D - 1) && (F_side != nullptr);
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
    for (int m = 0; m < n0; m++) {
        const int a = m + 1;
        const int s = S.sg[0][m];
        // neighbour's shell-s sums: from the phase-1 table if the perturbation touches it
        double tj, Tj;
        if (S.amask[a] & bit) {
            const int idx = (S.toff[a] + __popcll(S.amask[a] & (bit - 1ull))) * D + r;
            tj = S.ptj[idx];
            Tj = S.pTj[idx];
        } else {
            tj = S.bt[a][s];
            Tj = S.bT[a][s];
        }
        // owner's bond m
        double d, cx, cy, cz;
        if (own_all) {
            d = od[m];
            cx = ocx[m];
            cy = ocy[m];
            cz = ocz[m];
        } else if (m == mstar) {
            d = dn;
            cx = ncx;
            cy = ncy;
            cz = ncz;
        } else {
            d = S.bd[0][m];
            cx = S.bcs[m][0];
            cy = S.bcs[m][1];
            cz = S.bcs[m][2];
        }
        double f = 2.0 * S.Kn[m] * d + 0.5 * (Ti[s] + Tj) + 0.5 * S.Tv[0][m] * (ti[s] + tj);
        f *= S.brk[0][m];
        p0 += cx * f;
        p1 += cy * f;
        p2 += cz * f;
        if (last)
            F_side[(size_t)m * Npz + i] = f;
    }
    if (last) {
        Pin_side[i] = p0;
        Pin_side[Npz + i] = p1;
        Pin_side[2 * Npz + i] = p2;
    }
    // A_i[c][s][r] = (Pin_pert[s] - Pin_base[s]) / EPS / radius -> element (row s, col r) of block (i,c)
    const long long k = krow + q;
    const double pp[3] = {p0, p1, p2};
#pragma unroll
    for (int s = 0; s < D; s++) {
        const double kv = (pp[s] - S.base_pin[s]) / eps / radius;
        val[(k * D * D + s * D + r) * 32 + lane_i] = kv;
    }
}

// =====================================================================================================================
// Round-2 assembly for 3-D lattices: slice-cooperative, two kernels (param `fd_variant`, default 1; 0 = the kernel above).
//
// The CTA-per-particle kernel above spends 37 k warp instructions per particle: every owner re-derives the perturbed
// shell sums of its 18 neighbours (each neighbour's table is rebuilt by ~18 owners), stages 342 scattered bond records
// and writes K 8 bytes at a time into 256-byte lines.  The perturbed shell sums of a particle do not depend on who
// asks, so they are tabulated ONCE per particle, and the rows are then formed by CTAs that own a whole SELL slice
// (32 consecutive block rows, lane = row):
//
//  fd_shell_table_kernel  particle a -> planes of `tab` ([plane][Np], lane = particle: coalesced both ways)
//      0..3                       unperturbed dL_total[a][0..1], TdL_total[a][0..1]
//      4 + (mm*3 + r)*2 + {0,1}   shell sums (dL, TdL) of shell nsign[a][mm] when neighbour slot mm moves by +h in r
//      4 + NN*6 + r*4 + 2*s+{0,1} shell sums of shell s when a itself moves by +h in r
//    each sum is accumulated in the reference's order with exactly its expressions (constitutive.c:241-260); stretches
//    of bonds the perturbation does not touch are bit-identical to the unperturbed ones and are re-added from a cache.
//  fd_rows_kernel         one CTA per slice, warp w takes the conn positions q = w, w+8, ...; a THREAD owns the three
//      perturbations (conn[i][q], r = 0..2) of its row: a bond whose inputs the perturbation does not touch contributes
//      its cached unperturbed product cs*F (bit-identical: same inputs, same operations), the others are re-evaluated
//      with sums looked up in `tab`.  On a regular lattice the 32 rows of a slice share one topology pattern, so the
//      "is bond m touched by conn member q" branches are warp-uniform, and every store of K is a full 256-byte line.
//
// ~2 k warp instructions per particle instead of 37 k; the arithmetic per stored value is unchanged, so K_global / IK /
// JK stay bit-identical to stiffness.c:384-516 (tests/test_constitutive_gpu.py, test_variants_gpu.py and the
// old-vs-new comparison on perturbed, damaged lattices in tests/test_fd_variants_gpu.py).
// =====================================================================================================================

#define FD_TAB_PLANES(NN) (4 + (NN) * 6 + 12)

template <int NN>
__global__ void __launch_bounds__(32 * (NN + 1))
fd_shell_table_kernel(int N, int Np, int a0, int tabNp, double h, const int *__restrict__ nbr, const signed char *__restrict__ nsign, const int *__restrict__ nbi_g,
                      const double *__restrict__ xyz, const double *__restrict__ L0g, const double *__restrict__ dLp0g,
                      const double *__restrict__ brkg, const double *__restrict__ Tvg, double *__restrict__ tab)
{
    __shared__ double bd[NN][32], tv[NN][32], ds[3][NN][32];
    __shared__ signed char sgs[NN][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int a = a0 + blockIdx.x * 32 + lane;   // the table covers particles [a0, a0 + tabNp): window of the current chunk
    const size_t Npz = Np, Tz = tabNp;
    const bool valid = a < N;
    const int n = valid ? nbi_g[a] : 0;
    const bool have = w < NN && w < n;
    tab -= a0;
    double dn[3] = {0.0, 0.0, 0.0};
    int smm = 0;
    if (have) {
        const size_t g = (size_t)w * Npz + a;
        const int nj = nbr[g];
        const double pa[3] = {xyz[a], xyz[Npz + a], xyz[2 * Npz + a]};
        const double pn[3] = {xyz[nj], xyz[Npz + nj], xyz[2 * Npz + nj]};
        const double l0 = L0g[g], dlp = dLp0g[g], bk = brkg[g];
        smm = nsign[g];
        {
            const double dx = pa[0] - pn[0], dy = pa[1] - pn[1], dz = pa[2] - pn[2];
            const double dis = sqrt(dx * dx + dy * dy + dz * dz);
            double d = dis - l0;
            d -= dlp;
            d *= bk;
            bd[w][lane] = d;
        }
        tv[w][lane] = Tvg[g];
        sgs[w][lane] = (signed char)smm;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            // the neighbour moves (constitutive.c:248-250 with xyz[neighbors[a][w]][r] + h)
            double q[3] = {pn[0], pn[1], pn[2]};
            q[r] = q[r] + h;
            double dx = pa[0] - q[0], dy = pa[1] - q[1], dz = pa[2] - q[2];
            double dis = sqrt(dx * dx + dy * dy + dz * dz);
            double d = dis - l0;
            d -= dlp;
            d *= bk;
            dn[r] = d;
            // a itself moves
            double o[3] = {pa[0], pa[1], pa[2]};
            o[r] = o[r] + h;
            dx = o[0] - pn[0], dy = o[1] - pn[1], dz = o[2] - pn[2];
            dis = sqrt(dx * dx + dy * dy + dz * dz);
            d = dis - l0;
            d -= dlp;
            d *= bk;
            ds[r][w][lane] = d;
        }
    }
    __syncthreads();
    if (w < NN) {
        if (!have)
            return;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            double t = 0.0, T = 0.0;
            for (int m2 = 0; m2 < n; m2++) {
                if (sgs[m2][lane] != smm)
                    continue;
                const double d = m2 == w ? dn[r] : bd[m2][lane];
                t += d;
                T += tv[m2][lane] * d;
            }
            tab[(size_t)(4 + (w * 3 + r) * 2) * Tz + a] = t;
            tab[(size_t)(4 + (w * 3 + r) * 2 + 1) * Tz + a] = T;
        }
        return;
    }
    if (!valid)
        return;
    {
        double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
        for (int m2 = 0; m2 < n; m2++) {
            const double d = bd[m2][lane];
            const double td = tv[m2][lane] * d;
            if (sgs[m2][lane] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
        }
        tab[a] = t0;
        tab[Tz + a] = t1;
        tab[2 * Tz + a] = T0;
        tab[3 * Tz + a] = T1;
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
        for (int m2 = 0; m2 < n; m2++) {
            const double d = ds[r][m2][lane];
            const double td = tv[m2][lane] * d;
            if (sgs[m2][lane] == 0) {
                t0 += d;
                T0 += td;
            } else {
                t1 += d;
                T1 += td;
            }
        }
        double *o = tab + (size_t)(4 + NN * 6 + r * 4) * Tz + a;
        o[0] = t0;
        o[Tz] = T0;
        o[2 * Tz] = t1;
        o[3 * Tz] = T1;
    }
}

// x / y for a divisor used many times, given rcp = RN(1 / y), y > 0 and normal: two Newton corrections of the quotient
// with exact FMA residuals give the correctly rounded quotient (Markstein's theorem: a faithful quotient corrected once
// with the correctly rounded reciprocal is the IEEE quotient) as long as nothing under- or overflows -- 5 instructions
// instead of the ~25 (plus a slow-path call for zero numerators) of the generic fp64 division.  Numerators outside
// [2^-900, 2^900] (zero, subnormal, huge, Inf / NaN) take the real division.  Checked against `/` on 8.6e8 random pairs
// incl. the divisors used here (scripts/check_exact_division.c).
__device__ __noinline__ double fd_div_rare(double x, double y) { return x / y; }   // a call is never speculated
__device__ __forceinline__ double fd_div_by(double x, double y, double rcp)
{
    const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    if (ex - 123u < 1800u) {
        const double q0 = x * rcp;
        const double e0 = fma(-y, q0, x);
        const double q1 = fma(e0, rcp, q0);
        const double e1 = fma(-y, q1, x);
        return fma(e1, rcp, q1);
    }
    // zero numerators are common on an undeformed lattice (direction cosines, decoupled components); the generic
    // division would take its slow path for each of them
    if (x == 0.0)
        return x;
    return fd_div_rare(x, y);
}

template <int NN>
struct FdRowsSmem {
    int conn[64][32];                       // sorted conn list of every row of the slice
    int na[NN][32];                         // neighbour ids
    double bd[NN][32], cx[NN][32], cy[NN][32], cz[NN][32], Kn[NN][32], Tv[NN][32], brk[NN][32], fb[NN][32];
    double btj[NN][32], bTj[NN][32];        // unperturbed sums of neighbour m in the shell of bond m
    double bti[2][32], bTi[2][32];          // unperturbed sums of the row's own particle
    double bpin[3][32];
    unsigned long long amask[NN][32];       // bit q: conn member q changes neighbour m's sums in the shell of bond m
    unsigned long long ownmask[32];         // bit q: conn member q is one of the row's own neighbours
    unsigned char slotof[NN][NN + 1][32];   // rank of q inside amask -> slot in the neighbour's list (NN = the neighbour itself)
    unsigned char sg[NN][32];
    unsigned char ownq[NN][32];             // conn position of own neighbour m (255: none)
    unsigned char selfq[32];
};

// first position k in the sorted column `conn[.][lane]` (n entries) with conn[k] >= target
__device__ __forceinline__ int fd_lower_bound(const int (*conn)[32], int lane, int n, int target)
{
    int lo = 0, len = n;
    while (len > 0) {
        const int half = len >> 1;
        if (conn[lo + half][lane] < target) {
            lo += half + 1;
            len -= half + 1;
        } else
            len = half;
    }
    return lo;
}

template <int NN>
__global__ void __launch_bounds__(256, 2)
fd_rows_kernel(int N, int Np, int slice0, int a0, int tabNp, double h, double eps, double radius, double rcp_eps, double rcp_radius,
               const int *__restrict__ nbr, const signed char *__restrict__ nsign,
               const int *__restrict__ nbi_g, const double *__restrict__ xyz, const double *__restrict__ L0g,
               const double *__restrict__ dLp0g, const double *__restrict__ brkg, const double *__restrict__ Tvg,
               const double *__restrict__ Kng, const long long *__restrict__ sptr, const int *__restrict__ col,
               const int *__restrict__ nbc_g, const double *__restrict__ tab, double *__restrict__ val, double *__restrict__ F_side,
               double *__restrict__ Pin_side)
{
    extern __shared__ unsigned char smem_raw[];
    FdRowsSmem<NN> &S = *reinterpret_cast<FdRowsSmem<NN> *>(smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slice = slice0 + blockIdx.x;
    const int i = slice * 32 + lane;
    const bool valid = i < N;
    const size_t Npz = Np, Tz = tabNp;
    const long long krow = sptr[slice];
    const int width = (int)(sptr[slice + 1] - krow);
    const int n0 = valid ? nbi_g[i] : 0;
    const int nbc = valid ? nbc_g[i] : 0;
    tab -= a0;   // indexed by particle id; the window [a0, a0 + tabNp) covers the rows of this launch and their neighbours
    const double *tabN = tab + 4 * Tz, *tabS = tab + (size_t)(4 + NN * 6) * Tz;

    for (int k = w; k < width; k += 8)
        S.conn[k][lane] = col[(krow + k) * 32 + lane];
    if (w == 0)
        S.ownmask[lane] = 0ull;
    __syncthreads();

    // ---- the row's own bonds: unperturbed stretch and direction cosines, neighbour sums, conn positions ----
    double pi0 = 0.0, pi1 = 0.0, pi2 = 0.0;
    if (valid) {
        pi0 = xyz[i];
        pi1 = xyz[Npz + i];
        pi2 = xyz[2 * Npz + i];
    }
    for (int m = w; m < NN; m += 8) {
        if (m >= n0)
            continue;
        const size_t g = (size_t)m * Npz + i;
        const int a = nbr[g];
        const double dx = pi0 - xyz[a], dy = pi1 - xyz[Npz + a], dz = pi2 - xyz[2 * Npz + a];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        const double bk = brkg[g];
        double d = dis - L0g[g];
        d -= dLp0g[g];
        d *= bk;
        const int s = nsign[g];
        S.na[m][lane] = a;
        S.bd[m][lane] = d;
        const double rd = 1.0 / dis;
        S.cx[m][lane] = fd_div_by(dx, dis, rd);
        S.cy[m][lane] = fd_div_by(dy, dis, rd);
        S.cz[m][lane] = fd_div_by(dz, dis, rd);
        S.Kn[m][lane] = Kng[g];
        S.Tv[m][lane] = Tvg[g];
        S.brk[m][lane] = bk;
        S.sg[m][lane] = (unsigned char)s;
        S.btj[m][lane] = tab[(size_t)s * Tz + a];
        S.bTj[m][lane] = tab[(size_t)(2 + s) * Tz + a];
        const int pos = fd_lower_bound(S.conn, lane, nbc, a);
        const bool found = pos < nbc && S.conn[pos][lane] == a;
        S.ownq[m][lane] = found ? (unsigned char)pos : (unsigned char)255;
        if (found)
            atomicOr(&S.ownmask[lane], 1ull << pos);
    }
    if (w == 7) {
        unsigned char sq = 255;
        if (valid) {
            const int pos = fd_lower_bound(S.conn, lane, nbc, i);
            if (pos < nbc && S.conn[pos][lane] == i)
                sq = (unsigned char)pos;
            S.bti[0][lane] = tab[i];
            S.bti[1][lane] = tab[Tz + i];
            S.bTi[0][lane] = tab[2 * Tz + i];
            S.bTi[1][lane] = tab[3 * Tz + i];
        }
        S.selfq[lane] = sq;
    }
    __syncthreads();

    // ---- unperturbed bond forces (constitutive.c:269-270) and the "who touches whom" masks ----
    for (int m = w; m < NN; m += 8) {
        if (m >= n0)
            continue;
        const int s = S.sg[m][lane];
        double f = 2.0 * S.Kn[m][lane] * S.bd[m][lane] + 0.5 * (S.bTi[s][lane] + S.bTj[m][lane]) +
                   0.5 * S.Tv[m][lane] * (S.bti[s][lane] + S.btj[m][lane]);
        f *= S.brk[m][lane];
        S.fb[m][lane] = f;
        const int a = S.na[m][lane];
        const int nan_ = nbi_g[a];
        unsigned long long mask = 0ull;
        unsigned char qs[NN];
        int cc[NN];
        // the neighbour's bond list first (independent loads in flight together), then the searches
#pragma unroll
        for (int mm = 0; mm < NN; mm++) {
            cc[mm] = -1;
            if (mm < nan_) {
                const size_t gg = (size_t)mm * Npz + a;
                const int c = nbr[gg];
                if (nsign[gg] == s)
                    cc[mm] = c;
            }
        }
#pragma unroll
        for (int mm = 0; mm < NN; mm++) {
            qs[mm] = 255;
            if (cc[mm] >= 0) {
                const int pos = fd_lower_bound(S.conn, lane, nbc, cc[mm]);
                if (pos < nbc && S.conn[pos][lane] == cc[mm]) {
                    mask |= 1ull << pos;
                    qs[mm] = (unsigned char)pos;
                }
            }
        }
        const int oq = S.ownq[m][lane];
        if (oq != 255)
            mask |= 1ull << oq;
        S.amask[m][lane] = mask;
#pragma unroll
        for (int mm = 0; mm < NN; mm++)
            if (qs[mm] != 255)
                S.slotof[m][__popcll(mask & ((1ull << qs[mm]) - 1ull))][lane] = (unsigned char)mm;
        if (oq != 255)
            S.slotof[m][__popcll(mask & ((1ull << oq) - 1ull))][lane] = (unsigned char)NN;
    }
    __syncthreads();

    // ---- unperturbed internal force of every row (constitutive.c:275-277): warp 0 forms it while the others start on
    //      their first conn member; they meet at named barrier 1 before the first store of K ----
    if (w == 0 && valid) {
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        for (int m = 0; m < n0; m++) {
            const double f = S.fb[m][lane];
            q0 += S.cx[m][lane] * f;
            q1 += S.cy[m][lane] * f;
            q2 += S.cz[m][lane] * f;
        }
        S.bpin[0][lane] = q0;
        S.bpin[1][lane] = q1;
        S.bpin[2][lane] = q2;
    }
    bool met = false;

    // ---- one conn member per thread: the three perturbed evaluations of the row's internal force ----
    for (int q = w; q < width; q += 8) {
        const bool active = q < nbc;
        double p[3][3];
        bool last = false;
        if (active) {
        const unsigned long long bitq = 1ull << q, below = bitq - 1ull;
        const bool own_all = S.selfq[lane] == q;
        int mstar = -1, sstar = -1;
        if (!own_all && (S.ownmask[lane] & bitq)) {
            for (int m = 0; m < n0; m++)
                if (S.ownq[m][lane] == q)
                    mstar = m;
            sstar = S.sg[mstar][lane];
        }
        // the row's own shell sums under the three perturbations: [r][shell]
        double ot[3][2], oT[3][2];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            ot[r][0] = S.bti[0][lane];
            ot[r][1] = S.bti[1][lane];
            oT[r][0] = S.bTi[0][lane];
            oT[r][1] = S.bTi[1][lane];
        }
        if (own_all) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double *o = tabS + (size_t)(r * 4) * Tz + i;
                ot[r][0] = o[0];
                oT[r][0] = o[Tz];
                ot[r][1] = o[2 * Tz];
                oT[r][1] = o[3 * Tz];
            }
        } else if (mstar >= 0) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double t = tabN[(size_t)((mstar * 3 + r) * 2) * Tz + i], T = tabN[(size_t)((mstar * 3 + r) * 2 + 1) * Tz + i];
                if (sstar == 0) {
                    ot[r][0] = t;
                    oT[r][0] = T;
                } else {
                    ot[r][1] = t;
                    oT[r][1] = T;
                }
            }
        }
        last = (q == nbc - 1) && (F_side != nullptr);
#pragma unroll
        for (int r = 0; r < 3; r++)
            p[r][0] = p[r][1] = p[r][2] = 0.0;

        for (int m = 0; m < n0; m++) {
            const int s = S.sg[m][lane];
            const unsigned long long am = S.amask[m][lane];
            const bool affn = (am & bitq) != 0ull;
            const bool own_d = own_all || m == mstar;
            const bool own_s = own_all || s == sstar;
            if (!(affn || own_d || own_s)) {
                const double f = S.fb[m][lane];
                const double ax = S.cx[m][lane] * f, ay = S.cy[m][lane] * f, az = S.cz[m][lane] * f;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    p[r][0] += ax;
                    p[r][1] += ay;
                    p[r][2] += az;
                }
                if (last)
                    F_side[(size_t)m * Npz + i] = f;
                continue;
            }
            // neighbour's sums in the shell of bond m
            double tj[3], Tj[3];
            if (affn) {
                const int slot = S.slotof[m][__popcll(am & below)][lane];
                const int a = S.na[m][lane];
                if (slot == NN) {
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        tj[r] = tabS[(size_t)(r * 4 + 2 * s) * Tz + a];
                        Tj[r] = tabS[(size_t)(r * 4 + 2 * s + 1) * Tz + a];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        tj[r] = tabN[(size_t)((slot * 3 + r) * 2) * Tz + a];
                        Tj[r] = tabN[(size_t)((slot * 3 + r) * 2 + 1) * Tz + a];
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    tj[r] = S.btj[m][lane];
                    Tj[r] = S.bTj[m][lane];
                }
            }
            // the bond itself
            double d[3], ux[3], uy[3], uz[3];
            const double bk = S.brk[m][lane];
            if (own_d) {
                const size_t g = (size_t)m * Npz + i;
                const int a = S.na[m][lane];
                const double pa0 = xyz[a], pa1 = xyz[Npz + a], pa2 = xyz[2 * Npz + a];
                const double l0 = L0g[g], dlp = dLp0g[g];
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    double o[3] = {pi0, pi1, pi2}, e[3] = {pa0, pa1, pa2};
                    if (own_all)
                        o[r] = o[r] + h;
                    else
                        e[r] = e[r] + h;
                    const double dx = o[0] - e[0], dy = o[1] - e[1], dz = o[2] - e[2];
                    const double dis = sqrt(dx * dx + dy * dy + dz * dz);
                    double dd = dis - l0;
                    dd -= dlp;
                    dd *= bk;
                    d[r] = dd;
                    const double rd = 1.0 / dis;   // RN(1 / dis): three exact quotients for the price of one division
                    ux[r] = fd_div_by(dx, dis, rd);
                    uy[r] = fd_div_by(dy, dis, rd);
                    uz[r] = fd_div_by(dz, dis, rd);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    d[r] = S.bd[m][lane];
                    ux[r] = S.cx[m][lane];
                    uy[r] = S.cy[m][lane];
                    uz[r] = S.cz[m][lane];
                }
            }
            const double kn = S.Kn[m][lane], tvm = S.Tv[m][lane];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double ti = s ? ot[r][1] : ot[r][0], Ti = s ? oT[r][1] : oT[r][0];
                double f = 2.0 * kn * d[r] + 0.5 * (Ti + Tj[r]) + 0.5 * tvm * (ti + tj[r]);
                f *= bk;
                p[r][0] += ux[r] * f;
                p[r][1] += uy[r] * f;
                p[r][2] += uz[r] * f;
                if (last && r == 2)
                    F_side[(size_t)m * Npz + i] = f;
            }
        }
        }   // active
        if (!met) {
            __syncwarp();
            asm volatile("barrier.sync 1;" ::: "memory");
            met = true;
        }
        if (!active)
            continue;
        if (last) {
            Pin_side[i] = p[2][0];
            Pin_side[Npz + i] = p[2][1];
            Pin_side[2 * Npz + i] = p[2][2];
        }
        // A_i[c][s][r] = (Pin_pert[s] - Pin_base[s]) / EPS / radius -> element (row s, col r) of block (i, c)
        const long long k = krow + q;
#pragma unroll
        for (int s = 0; s < 3; s++) {
            const double b = S.bpin[s][lane];
#pragma unroll
            for (int r = 0; r < 3; r++)
                val[(k * 9 + s * 3 + r) * 32 + lane] = fd_div_by(fd_div_by(p[r][s] - b, eps, rcp_eps), radius, rcp_radius);
        }
    }
    if (!met) {
        __syncwarp();
        asm volatile("barrier.sync 1;" ::: "memory");
    }
}

// position of block column `target` in (sorted) block row `row`, or -1
__device__ __forceinline__ int find_col(const int *__restrict__ col, const long long *__restrict__ sptr, const int *__restrict__ nbc, int row,
                                        int target)
{
    const long long base = sptr[row >> 5] * 32 + (row & 31);
    int lo = 0, hi = nbc[row] - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = col[base + (long long)mid * 32];
        if (v == target)
            return mid;
        if (v < target)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    return -1;
}

// K_ij = 0.5*A_i[j] + 0.5*A_j[i]^T for j > i (both copies written); diagonal block: lower := upper.
template <int D>
__global__ void symmetrize_kernel(int N, const long long *__restrict__ sptr, const int *__restrict__ col, const int *__restrict__ nbc,
                                  double *__restrict__ val)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N)
        return;
    const long long ka = sptr[row >> 5];
    const int lane = row & 31;
    const int n = nbc[row];
    for (int kk = 0; kk < n; kk++) {
        const long long k = ka + kk;
        const int cidx = col[k * 32 + lane];
        if (cidx < row)
            continue;
        double *a = val + (k * D * D) * 32 + lane;
        if (cidx == row) {
#pragma unroll
            for (int r = 1; r < D; r++)
#pragma unroll
                for (int s = 0; s < r; s++)
                    a[(r * D + s) * 32] = a[(s * D + r) * 32];
            continue;
        }
        const int pos = find_col(col, sptr, nbc, cidx, row);
        if (pos < 0)
            continue;  // asymmetric pattern: cannot happen for conn built by neighbor.c
        double *b = val + ((sptr[cidx >> 5] + pos) * D * D) * 32 + (cidx & 31);
        // read both blocks completely first, then write (in-place safe)
        double A[D * D], B[D * D];
#pragma unroll
        for (int e = 0; e < D * D; e++) {
            A[e] = a[e * 32];
            B[e] = b[e * 32];
        }
#pragma unroll
        for (int r = 0; r < D; r++)
#pragma unroll
            for (int s = 0; s < D; s++) {
                const double kv = 0.5 * A[r * D + s] + 0.5 * B[s * D + r];
                a[(r * D + s) * 32] = kv;
                b[(s * D + r) * 32] = kv;
            }
    }
}

// after the (single-threaded) reference assembly, dL / cs* / dL_total / TdL_total of
// particle p hold the values of the LAST evaluation that touched p: the evaluation of particle
// m(p) = max({p} U {intact neighbours whose own bond to p is intact}) with its last conn particle
// displaced by +h in the last coordinate.
template <int D>
__global__ void __launch_bounds__(128)
fd_side_effects_kernel(int N, int Np, double h, const int *__restrict__ nbr, const signed char *__restrict__ nsign,
                       const signed char *__restrict__ mirror, const int *__restrict__ nbi_g, const double *__restrict__ xyz,
                       const double *__restrict__ L0, const double *__restrict__ dLp0, const double *__restrict__ brk, const double *__restrict__ Tv,
                       const long long *__restrict__ sptr, const int *__restrict__ col, const int *__restrict__ nbc, double *__restrict__ dL,
                       double *__restrict__ csx, double *__restrict__ csy, double *__restrict__ csz, double *__restrict__ dLt, double *__restrict__ TdLt)
{
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= N)
        return;
    const size_t Npz = Np;
    const int n = nbi_g[p];
    int m = p;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Npz + p;
        const int nj = nbr[e];
        const int mj = mirror[e];
        // p is in the star of nj iff nj's own bond to p is intact (constitutive.c:233-237)
        if (mj >= 0 && brk[(size_t)mj * Npz + nj] > LPMB_EPS && nj > m)
            m = nj;
    }
    if (nbc[m] <= 0)
        return;
    const int c = col[(sptr[m >> 5] + nbc[m] - 1) * 32 + (m & 31)];
    double pp[3] = {xyz[p], xyz[Npz + p], xyz[2 * Npz + p]};
    if (p == c)
        pp[D - 1] = pp[D - 1] + h;
    double t0 = 0, t1 = 0, T0 = 0, T1 = 0;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Npz + p;
        const int nj = nbr[e];
        double np[3] = {xyz[nj], xyz[Npz + nj], xyz[2 * Npz + nj]};
        if (nj == c)
            np[D - 1] = np[D - 1] + h;
        const double dx = pp[0] - np[0], dy = pp[1] - np[1], dz = pp[2] - np[2];
        const double dis = sqrt(dx * dx + dy * dy + dz * dz);
        double d = dis - L0[e];
        d -= dLp0[e];
        d *= brk[e];
        dL[e] = d;
        const double td = Tv[e] * d;
        if (nsign[e] == 0) {
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
    dLt[p] = t0;
    dLt[Npz + p] = t1;
    TdLt[p] = T0;
    TdLt[Npz + p] = T1;
}

// max |neighbour id - own id| over all bonds (one int back to the host; 0.2 ms at 10 M particles)
__global__ void fd_bandwidth_kernel(int N, int Np, int nn, const int *__restrict__ nbr, const int *__restrict__ nbi, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0;
    if (i < N) {
        const int n = nbi[i];
        for (int m = 0; m < n && m < nn; m++) {
            const int d = nbr[(size_t)m * Np + i] - i;
            b = max(b, d < 0 ? -d : d);
        }
    }
    for (int o = 16; o > 0; o >>= 1)
        b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b > 0)
        atomicMax(out, b);
}

static int fd_bandwidth(lpmb_ctx *c, const int *nbr, const int *nbi, int *band)
{
    LPMB_TRY(lpmb_ensure_staging(c, 64));
    int *d = (int *)c->staging;
    LPMB_MEMSET(c, d, 0, sizeof(int));
    fd_bandwidth_kernel<<<lpmb_blocks(c->N, 256), 256, 0, c->stream>>>(c->N, c->Np, c->nn, nbr, nbi, d);
    LPMB_LAUNCH_CHECK(c);
    LPMB_D2H(c, band, d, sizeof(int));
    return LPMB_OK;
}

extern "C" int lpmb_fd_stiffness(lpmb_ctx *c, int emulate_side_effects)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_REQUIRE(c->params.count("radius"), LPMB_ERR_STATE, "parameter radius not set");
    LPMB_TRY(lpmb_matrix_alloc_values(c));
    SellMatrix &K = c->K;
    const double radius = param(c, "radius");
    const double eps = LPMB_EPS;
    const double h = eps * radius;  // stiffness.c:419: xtemp + EPS * radius
    const int *nbr = fptr<int>(c, "neighbors");
    const signed char *nsign = fptr<signed char>(c, "nsign"), *mirror = fptr<signed char>(c, "mirror");
    const int *nbi = fptr<int>(c, "nb_initial");
    const double *xyz = fptr<double>(c, "xyz");
    const double *L0 = fptr<double>(c, "distance_initial"), *dLp0 = fptr<double>(c, "dLp0"), *brk = fptr<double>(c, "damage_broken");
    const double *Tv = fptr<double>(c, "Tv"), *Kn = fptr<double>(c, "Kn");
    double *F = fptr<double>(c, "F"), *Pin = fptr<double>(c, "Pin");
    LPMB_REQUIRE(nbr && nsign && mirror && nbi && xyz && L0 && dLp0 && brk && Tv && Kn && F && Pin, LPMB_ERR_STATE, "fields missing");
    double *Fs = emulate_side_effects ? F : nullptr, *Ps = emulate_side_effects ? Pin : nullptr;
    if (c->dim == 3) {
        LPMB_REQUIRE(c->nn <= 18 && c->nconn <= 64, LPMB_ERR_UNSUPPORTED, "fd_stiffness<3>: nn=%d nconn=%d", c->nn, c->nconn);
        // slice-cooperative assembly (default).  Its table of perturbed shell sums is persistent scratch of bounded size
        // (param fd_tab_mb, default 2048): the rows are processed in chunks of consecutive SELL slices and the table
        // covers the chunk's rows plus `band` = max |neighbour id - own id| particles either side (x-fastest lattices:
        // one or two lattice layers).  fd_variant = 0 -- or a numbering so scattered that no chunk fits the budget --
        // selects the CTA-per-particle kernel (same bits).
        bool done = false;
        if ((int)param(c, "fd_variant", 1.0) != 0) {
            int band = 0;
            LPMB_TRY(fd_bandwidth(c, nbr, nbi, &band));
            const size_t per_particle = (size_t)FD_TAB_PLANES(18) * sizeof(double);
            const size_t budget = (size_t)(param(c, "fd_tab_mb", 2048.0) * 1048576.0);
            long long W = (long long)(budget / per_particle) & ~31ll;   // particles the table may hold
            if (W > c->Np)
                W = c->Np;
            const long long band32 = ((long long)band + 31) & ~31ll;
            long long rows = W >= c->Np ? c->Np : ((W - 2 * band32) & ~31ll);  // rows per chunk
            if (rows >= 256 || rows >= c->Np) {
                if (c->fd_tab_bytes < (size_t)W * per_particle) {
                    LPMB_CUDA(cudaStreamSynchronize(c->stream));
                    cudaFree(c->fd_tab);
                    c->fd_tab = nullptr;
                    c->fd_tab_bytes = 0;
                    if (cudaMalloc(&c->fd_tab, (size_t)W * per_particle) == cudaSuccess)
                        c->fd_tab_bytes = (size_t)W * per_particle;
                    else
                        (void)cudaGetLastError();
                }
                if (c->fd_tab) {
                    const size_t smem = sizeof(FdRowsSmem<18>);
                    auto kern = fd_rows_kernel<18>;
                    LPMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    for (long long p0 = 0; p0 < c->Np; p0 += rows) {
                        const long long p1 = p0 + rows < c->Np ? p0 + rows : c->Np;
                        long long a0 = p0 - band32, a1 = p1 + band32;
                        a0 = a0 < 0 ? 0 : a0;
                        a1 = a1 > c->Np ? c->Np : a1;
                        const int tabNp = (int)(a1 - a0);
                        fd_shell_table_kernel<18><<<tabNp / 32, 32 * 19, 0, c->stream>>>(c->N, c->Np, (int)a0, tabNp, h, nbr, nsign, nbi, xyz, L0, dLp0, brk, Tv, c->fd_tab);
                        LPMB_LAUNCH_CHECK(c);
                        kern<<<(int)((p1 - p0) / 32), 256, smem, c->stream>>>(c->N, c->Np, (int)(p0 / 32), (int)a0, tabNp, h, eps, radius, 1.0 / eps, 1.0 / radius, nbr, nsign, nbi, xyz, L0, dLp0, brk, Tv, Kn, K.sptr, K.col, K.nbc, c->fd_tab, K.val, Fs, Ps);
                        LPMB_LAUNCH_CHECK(c);
                    }
                    done = true;
                }
            }
        }
        if (!done) {
            const size_t smem = sizeof(StarSmem<3, 18>);
            auto kern = fd_stiffness_kernel<3, 18, 192>;
            LPMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<c->N, 192, smem, c->stream>>>(c->N, c->Np, h, eps, radius, nbr, nsign, nbi, xyz, L0, dLp0, brk, Tv, Kn, K.sptr, K.col, K.nbc, K.val, Fs, Ps);
            LPMB_LAUNCH_CHECK(c);
        }
        symmetrize_kernel<3><<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, K.sptr, K.col, K.nbc, K.val);
        LPMB_LAUNCH_CHECK(c);
    } else {
        LPMB_REQUIRE(c->nn <= 12 && c->nconn <= 32, LPMB_ERR_UNSUPPORTED, "fd_stiffness<2>: nn=%d nconn=%d", c->nn, c->nconn);
        const size_t smem = sizeof(StarSmem<2, 12>);
        auto kern = fd_stiffness_kernel<2, 12, 64>;
        LPMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<c->N, 64, smem, c->stream>>>(c->N, c->Np, h, eps, radius, nbr, nsign, nbi, xyz, L0, dLp0, brk, Tv, Kn, K.sptr, K.col, K.nbc, K.val, Fs, Ps);
        LPMB_LAUNCH_CHECK(c);
        symmetrize_kernel<2><<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, K.sptr, K.col, K.nbc, K.val);
        LPMB_LAUNCH_CHECK(c);
    }
    K.values_ready = true;
    lpmb_brick_touch(c);
    if (emulate_side_effects) {
        double *dL = fptr<double>(c, "dL"), *csx = fptr<double>(c, "csx"), *csy = fptr<double>(c, "csy"), *csz = fptr<double>(c, "csz");
        double *dLt = fptr<double>(c, "dL_total"), *TdLt = fptr<double>(c, "TdL_total");
        if (c->dim == 3)
            fd_side_effects_kernel<3><<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, c->Np, h, nbr, nsign, mirror, nbi, xyz, L0, dLp0, brk, Tv, K.sptr, K.col, K.nbc, dL, csx, csy, csz, dLt, TdLt);
        else
            fd_side_effects_kernel<2><<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, c->Np, h, nbr, nsign, mirror, nbi, xyz, L0, dLp0, brk, Tv, K.sptr, K.col, K.nbc, dL, csx, csy, csz, dLt, TdLt);
        LPMB_LAUNCH_CHECK(c);
    }
    return LPMB_OK;
}
