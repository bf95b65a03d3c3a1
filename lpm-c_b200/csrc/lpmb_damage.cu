// lpmb_damage.cu -- damage accumulation and bond breaking (compiled -fmad=false).
//
// Replaces, in the reference:
//   updateDamageGeneral(file, step, plmode)       src/constitutive.c:149-164
//   updateDuctileDamagePwiseNonlocal(file, step)  src/constitutive.c:1757-1862  (plmode 0)
//   updateBrittleDamage(file, step, nbreak)       src/constitutive.c:1437-1526  (plmode 6)
//   updateDuctileDamageBwiseLocal(file, step)     src/constitutive.c:1607-1695  (plmode 5)
//   updateDuctileDamagePwiseLocal(file, step)     src/constitutive.c:1529-1579  (LPMB_DAMAGE_PWISE_LOCAL; commented out in the dispatcher)
//   updateDuctileDamageBwiseNonlocal(file, step)  src/constitutive.c:1698-1753  (LPMB_DAMAGE_BWISE_NONLOCAL; likewise)
//
// Nonlocal law: D_i += sum_j phi(d_ij) V Ddot_j / sum_j phi(d_ij) V over ALL particles with
// d_ij < 3*damage_L in the initial configuration (self included).  The reference does this as an
// O(N^2) all-pairs loop; here the candidates come from the uniform cell grid (lpmb_topology.cu) built
// over xyz_initial with cell size 3*damage_L, i.e. 27 cells per particle.  The Gaussian uses exp(), so
// this kernel is compared at 1e-12 relative, not bit for bit.
#include <algorithm>

#include "lpmb_internal.cuh"

#define DT 128
#define LPMB_PI 3.14159265358979323846

__global__ void __launch_bounds__(DT)
nonlocal_damage_kernel(int N, int Np, CellGrid g, const double *__restrict__ x0 /* xyz_initial */, double L, double thr, double Ac, double V,
                       const double *__restrict__ dlambda, const double *__restrict__ triax, double *__restrict__ Dn /* damage_nonlocal[.][0] */)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    double Di = Dn[i];
    if (Di > thr) {  // constitutive.c:1802-1807: frozen, clamped
        if (Di > 1.0)
            Dn[i] = 1.0;
        return;
    }
    const double xi = x0[i], yi = x0[(size_t)Np + i], zi = x0[(size_t)2 * Np + i];
    const double inv = 1.0 / g.cell;
    int cx = (int)floor((xi - g.ox) * inv), cy = (int)floor((yi - g.oy) * inv), cz = (int)floor((zi - g.oz) * inv);
    cx = cx < 0 ? 0 : (cx >= g.nx ? g.nx - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= g.ny ? g.ny - 1 : cy);
    cz = cz < 0 ? 0 : (cz >= g.nz ? g.nz - 1 : cz);
    const double cut = 3 * L;
    double Ddot = 0, A = 0;
    for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= g.nz)
            continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = cy + dy;
            if (y < 0 || y >= g.ny)
                continue;
            for (int dx = -1; dx <= 1; dx++) {
                const int x = cx + dx;
                if (x < 0 || x >= g.nx)
                    continue;
                const int cidx = x + g.nx * (y + g.ny * z);
                for (int t = g.start[cidx]; t < g.start[cidx + 1]; t++) {
                    const int j = g.items[t];
                    const double ax = x0[j] - xi, ay = x0[(size_t)Np + j] - yi, az = x0[(size_t)2 * Np + j] - zi;
                    const double dis = sqrt(ax * ax + ay * ay + az * az);
                    if (dis < cut) {
                        double DdotLocal = 0;
                        const double f = (1.0 + Ac * triax[j]);
                        if (f > 0.0)
                            DdotLocal = dlambda[j] * (1.0 + Ac * triax[j]);
                        // DAM_PHI(x) = 1.0 / damage_L / sqrt(2*PI) * exp(-0.5*x*x / damage_L / damage_L)   (lpm.h:51)
                        const double phi = 1.0 / L / sqrt(2 * LPMB_PI) * exp(-0.5 * dis * dis / L / L);
                        Ddot += DdotLocal * phi * V;
                        A += phi * V;
                    }
                }
            }
        }
    }
    if (Ddot > 0.0)
        Dn[i] = Di + 1.0 / A * Ddot;
}

// constitutive.c:1828-1858 fused per bond: break if either end is beyond the threshold, then
// damage_D = max(D_i, D_j) on intact bonds and damage_w = 1 - damage_D.
__global__ void __launch_bounds__(DT)
nonlocal_break_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, double thr, const double *__restrict__ Dn,
                      double *__restrict__ broken, double *__restrict__ dD0, double *__restrict__ w, signed char *__restrict__ newly,
                      int *__restrict__ count, int own0, int own1)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const bool owned = i >= own0 && i < own1;  // ghost bonds (multi-GPU) are updated too but counted by their owner
    const double Di = Dn[i];
    const int n = nbi[i];
    int k = 0;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double Dj = Dn[nbr[e]];
        double b = broken[e], d = dD0[e];
        signed char nw = 0;
        if (Di > thr || Dj > thr) {
            if (fabs(b) > LPMB_EPS) {
                b = 0.0;
                d = 1.0;
                nw = 1;
                k++;
            }
        }
        if (fabs(b) > LPMB_EPS)
            d = Di < Dj ? Dj : Di;  // MAX(x,y) ((x) < (y) ? (y) : (x))
        broken[e] = b;
        dD0[e] = d;
        w[e] = 1.0 - d;
        newly[e] = nw;
    }
    if (k && owned)
        atomicAdd(count, k);
}

// ---- local bond-wise ductile damage (plmode 5), updateDuctileDamageBwiseLocal, constitutive.c:1607-1695 ------------
// (1) damage_local += (1 + A triax) dlambda below the threshold (clamped to 1 above it)
__global__ void __launch_bounds__(DT)
local_damage_kernel(int N, double thr, double Ac, const double *__restrict__ triax, const double *__restrict__ dlambda, double *__restrict__ dloc)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double f = (1.0 + Ac * triax[i]);
    double d = dloc[i];
    if (f > 0.0 && d <= thr)
        d += f * dlambda[i];
    else if (d > 1.0)
        d = 1.0;
    dloc[i] = d;
}

// (2) per bond D = (d_i + d_j)/2; an intact bond beyond the threshold breaks (the serial loop marks both directions
// when it meets the first one, so a bond is counted and logged once, from its lower-numbered end); nb is recounted
__global__ void __launch_bounds__(DT)
local_break_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, double thr, const double *__restrict__ dloc,
                   double *__restrict__ broken, double *__restrict__ dD0, int *__restrict__ nb, signed char *__restrict__ newly,
                   int *__restrict__ count)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double di = dloc[i];
    const int n = nbi[i];
    int k = 0, left = n;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = nbr[e];
        double D = 0.5 * (di + dloc[nj]);
        double b = broken[e];
        signed char nw = 0;
        if (D > thr && b > LPMB_EPS) {
            D = 1.0;
            b = 0.0;
            if (i < nj) {
                nw = 1;
                k++;
            }
        }
        if (b <= LPMB_EPS)
            left--;
        broken[e] = b;
        dD0[e] = D;
        newly[e] = nw;
    }
    nb[i] = left;
    if (k)
        atomicAdd(count, k);
}

// (3) broken bonds and bonds of fully detached particles carry D = 1; damage_w = 1 - D
__global__ void __launch_bounds__(DT)
local_finish_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, const int *__restrict__ nb,
                    const double *__restrict__ broken, double *__restrict__ dD0, double *__restrict__ w)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const int n = nbi[i], nbi_cur = nb[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        double D = dD0[e];
        if (fabs(broken[e]) < LPMB_EPS || nbi_cur == 0 || nb[nbr[e]] == 0)
            D = 1.0;
        dD0[e] = D;
        w[e] = 1.0 - D;
    }
}

// ---- particle-wise local ductile damage, updateDuctileDamagePwiseLocal, constitutive.c:1529-1579 -------------------
// (1) damage_local += (1 + A triax) dlambda at or below the threshold; a particle that is then beyond it (and not
// already at 1 within EPS) is set to 1: "newly detached".  The serial loop only reads particle i's own values here.
__global__ void __launch_bounds__(DT)
pwise_local_damage_kernel(int N, double thr, double Ac, const double *__restrict__ triax, const double *__restrict__ dlambda,
                          double *__restrict__ dloc, signed char *__restrict__ newp, int *__restrict__ count)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double f = (1.0 + Ac * triax[i]);
    double d = dloc[i];
    if (f > 0.0 && d <= thr)
        d += f * dlambda[i];
    signed char nw = 0;
    if (d > thr && fabs(d - 1.0) > LPMB_EPS) {
        d = 1.0;
        nw = 1;
        atomicAdd(count, 1);
    }
    dloc[i] = d;
    newp[i] = nw;
}

// (2) a newly detached particle zeroes damage_broken on all its bonds and on every slot of a neighbour that lists it
// (constitutive.c:1551-1565; `mirror` >= 0 says the neighbour does); then damage_D = MAX of the end values, w = 1 - D
__global__ void __launch_bounds__(DT)
pwise_local_break_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, const signed char *__restrict__ mirror,
                         const double *__restrict__ dloc, const signed char *__restrict__ newp, double *__restrict__ broken,
                         double *__restrict__ dD0, double *__restrict__ w)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double di = dloc[i];
    const bool mine = newp[i] != 0;
    const int n = nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = nbr[e];
        if (mine || (newp[nj] && mirror[e] >= 0))
            broken[e] = 0.0;
        const double dj = dloc[nj];
        const double D = di < dj ? dj : di;  // MAX(x,y) ((x) < (y) ? (y) : (x))
        dD0[e] = D;
        w[e] = 1.0 - D;
    }
}

// ---- bond-wise nonlocal ductile damage, updateDuctileDamageBwiseNonlocal, constitutive.c:1698-1753 ------------------
// (1) Gaussian average over the particle's own bond list: the particle itself enters with weight V (no phi), a
// neighbour with DAM_PHI(distance_initial) V; frozen / clamped beyond the threshold.  exp() -> compared at 1e-12.
__global__ void __launch_bounds__(DT)
bwise_nonlocal_damage_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, const double *__restrict__ L0, double L,
                             double thr, double Ac, double V, const double *__restrict__ dlambda, const double *__restrict__ triax,
                             double *__restrict__ Dn)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double Di = Dn[i];
    if (Di > thr) {
        if (Di > 1.0)
            Dn[i] = 1.0;
        return;
    }
    double DdotLocal = 0;
    double f = (1.0 + Ac * triax[i]);
    if (f > 0.0)
        DdotLocal = dlambda[i] * (1.0 + Ac * triax[i]);
    double Ddot = DdotLocal * V;
    double A = V;
    const int n = nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const int nj = nbr[e];
        DdotLocal = 0;
        f = (1.0 + Ac * triax[nj]);
        if (f > 0.0)
            DdotLocal = dlambda[nj] * (1.0 + Ac * triax[nj]);
        const double x = L0[e];
        const double phi = 1.0 / L / sqrt(2 * LPMB_PI) * exp(-0.5 * x * x / L / L);
        Ddot += DdotLocal * phi * V;
        A += phi * V;
    }
    if (Ddot > 0.0)
        Dn[i] = Di + 1.0 / A * Ddot;
}

// (2)+(3) fused per bond: an intact bond whose mean end damage is beyond the threshold breaks (D = 1; every direction
// is counted on its own, as the reference's double loop does); intact bonds carry the mean; w = 1 - D
__global__ void __launch_bounds__(DT)
bwise_nonlocal_break_kernel(int N, int Np, const int *__restrict__ nbi, const int *__restrict__ nbr, double thr, const double *__restrict__ Dn,
                            double *__restrict__ broken, double *__restrict__ dD0, double *__restrict__ w, signed char *__restrict__ newly,
                            int *__restrict__ count)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N)
        return;
    const double Di = Dn[i];
    const int n = nbi[i];
    int k = 0;
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double m = 0.5 * (Di + Dn[nbr[e]]);
        double b = broken[e], d = dD0[e];
        signed char nw = 0;
        if (m > thr && fabs(b) > LPMB_EPS) {
            b = 0.0;
            d = 1.0;
            nw = 1;
            k++;
        }
        if (fabs(b) > LPMB_EPS)
            d = m;
        broken[e] = b;
        dD0[e] = d;
        w[e] = 1.0 - d;
        newly[e] = nw;
    }
    if (k)
        atomicAdd(count, k);
}

// brittle: candidates with dL/L0 >= critical_bstrain, appended to a device list (order restored on the host)
__global__ void __launch_bounds__(DT)
brittle_candidates_kernel(int N, int Np, int nn, const int *__restrict__ nbi, const double *__restrict__ dL, const double *__restrict__ L0,
                          double crit, int cap, int *__restrict__ count, int *__restrict__ keys, double *__restrict__ strain, int own0, int own1)
{
    const int i = blockIdx.x * DT + threadIdx.x;
    if (i >= N || i < own0 || i >= own1)   // slab runs: every bond is a candidate on the rank that owns its particle
        return;
    const int n = nbi[i];
    for (int j = 0; j < n; j++) {
        const size_t e = (size_t)j * Np + i;
        const double s = dL[e] / L0[e];
        if (s >= crit) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) {
                keys[slot] = i * nn + j;
                strain[slot] = s;
            }
        }
    }
}

__global__ void brittle_apply_kernel(int nbreak, int nn, int Np, const int *__restrict__ keys, double *__restrict__ broken, double *__restrict__ dD0,
                                     double *__restrict__ w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbreak)
        return;
    const int i = keys[t] / nn, j = keys[t] % nn;
    const size_t e = (size_t)j * Np + i;
    dD0[e] = 1.0;
    w[e] = 0.0;
    broken[e] = 0.0;
}

// The selection of updateBrittleDamage (constitutive.c:1489-1520) on a candidate list in the reference's scan order
// (particle, then slot ascending == key ascending): all k candidates break if k <= nbreak, else the nbreak largest strains
// after the reference's shell sort -- reproduced verbatim in behaviour (it is not stable: ties are resolved exactly as
// there).  Sorts keys / strains in place; the bonds to break are [first, k).  Pure host arithmetic: in slab runs every rank
// runs it on the same all-gathered list (tests/test_oracle_port.py checks it against the oracle's restatement, also with
// the list assembled from per-rank pieces).
extern "C" int lpmb_brittle_select(int k, long long *keys, double *strains, int nbreak, int *first)
{
    LPMB_REQUIRE(k >= 0 && first && (k == 0 || (keys && strains)), LPMB_ERR_ARG, "lpmb_brittle_select: bad argument");
    *first = 0;
    if (k > nbreak) {
        for (int r = k / 2; r >= 1; r = r / 2)
            for (int a2 = r; a2 < k; ++a2) {
                const long long ti = keys[a2];
                const double tb = strains[a2];
                int b2 = a2 - r;
                while (b2 >= 0 && strains[b2] > tb) {
                    strains[b2 + r] = strains[b2];
                    keys[b2 + r] = keys[b2];
                    b2 = b2 - r;
                }
                strains[b2 + r] = tb;
                keys[b2 + r] = ti;
            }
        *first = k - (nbreak > 0 ? nbreak : 0);
    }
    return LPMB_OK;
}

extern "C" int lpmb_update_damage(lpmb_ctx *c, int plmode, int *broken_out, int *pairs, int max_pairs)
{
    LPMB_REQUIRE(c && broken_out, LPMB_ERR_ARG, "lpmb_update_damage: null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    *broken_out = 0;
    const int N = c->N, Np = c->Np, nn = c->nn;
    const int g = lpmb_blocks(N, DT);
    int *nbi = fptr<int>(c, "nb_initial"), *nbr = fptr<int>(c, "neighbors");
    double *broken = fptr<double>(c, "damage_broken"), *dD0 = fptr<double>(c, "damage_D0"), *w = fptr<double>(c, "damage_w");
    LPMB_REQUIRE(nbi && nbr && broken && dD0 && w, LPMB_ERR_STATE, "fields missing");
    int *d_count;
    LPMB_CUDA(cudaMalloc(&d_count, sizeof(int)));
    LPMB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
    int rc = LPMB_OK;
    if (plmode == 0) {
        for (const char *p : {"damage_L", "damage_threshold", "damagec_A", "particle_volume"})
            if (!c->params.count(p)) {
                cudaFree(d_count);
                lpmb_set_error("parameter %s not set", p);
                return LPMB_ERR_STATE;
            }
        const double L = param(c, "damage_L"), thr = param(c, "damage_threshold");
        CellGrid *grid = nullptr;
        const double *x0 = fptr<double>(c, "xyz_initial");
        // (re)built every call (once per load step): the topology builder shares the per-context grid slot
        rc = lpmb_grid_build(c, x0, 3 * L * 1.0000001, &grid);
        if (rc != LPMB_OK) {
            cudaFree(d_count);
            return rc;
        }
        // multi-GPU: the Gaussian reaches 3*damage_L beyond the owned slab -> refresh the ghosts' inputs first
        LPMB_TRY(lpmb_dist_exchange(c, fptr<double>(c, "J2_dlambda"), 1, true));
        LPMB_TRY(lpmb_dist_exchange(c, fptr<double>(c, "J2_triaxiality"), 1, true));
        nonlocal_damage_kernel<<<g, DT, 0, c->stream>>>(N, Np, *grid, x0, L, thr, param(c, "damagec_A"), param(c, "particle_volume"),
                                                        fptr<double>(c, "J2_dlambda"), fptr<double>(c, "J2_triaxiality"),
                                                        fptr<double>(c, "damage_nonlocal0"));
        LPMB_LAUNCH_CHECK(c);
        signed char *newly = nullptr;
        LPMB_CUDA(cudaMalloc(&newly, (size_t)nn * Np));
        // ... and the ghosts' damage itself (their own Gaussian would reach beyond the local block)
        LPMB_TRY(lpmb_dist_exchange(c, fptr<double>(c, "damage_nonlocal0"), 1, true));
        nonlocal_break_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, thr, fptr<double>(c, "damage_nonlocal0"), broken, dD0, w, newly, d_count,
                                                       lpmb_own0(c), lpmb_own1(c));
        LPMB_LAUNCH_CHECK(c);
        int k = 0;
        LPMB_CUDA(cudaMemcpyAsync(&k, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->world > 1) {  // global count of newly broken bonds
            LPMB_TRY(lpmb_cg_alloc(c));
            double kd = (double)k;
            LPMB_CUDA(cudaMemcpyAsync(c->cg.scal + 9, &kd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            LPMB_TRY(lpmb_dist_allreduce_sum(c, c->cg.scal + 9, 1));
            LPMB_CUDA(cudaMemcpyAsync(&kd, c->cg.scal + 9, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            k = (int)(kd + 0.5);
        }
        *broken_out = k;
        if (k > 0 && pairs && max_pairs > 0) {
            // reference logging order: i ascending, then slot j ascending (constitutive.c:1829-1845)
            std::vector<signed char> hn((size_t)nn * Np);
            std::vector<int> hnbr((size_t)nn * Np);
            LPMB_D2H(c, hn.data(), newly, hn.size());
            LPMB_D2H(c, hnbr.data(), nbr, hnbr.size() * sizeof(int));
            int t = 0;
            for (int i = 0; i < N && t < max_pairs; i++)
                for (int j = 0; j < nn && t < max_pairs; j++)
                    if (hn[(size_t)j * Np + i]) {
                        pairs[2 * t] = i;
                        pairs[2 * t + 1] = hnbr[(size_t)j * Np + i];
                        t++;
                    }
        }
        cudaFree(newly);
    } else if (plmode == 6) {
        // Slab runs (SURVEY 8(e) / constitutive.c:1489-1520): the nbreak largest strains are selected GLOBALLY.  Every rank lists
        // the candidates among the bonds of its owned particles, the lists travel to all ranks (two all-gathers: sizes, then
        // keys + strains padded to the longest list), and every rank runs the reference's selection on the same global list --
        // rank order = ascending global particle index, so the list is in the reference's scan order -- and applies the winners
        // that fall on its owned particles or on its ghosts with complete stars.
        LPMB_REQUIRE(c->params.count("critical_bstrain") && c->params.count("nbreak"), LPMB_ERR_STATE, "critical_bstrain / nbreak not set");
        const int cap = 1 << 16;
        int *keys;
        double *strain;
        LPMB_CUDA(cudaMalloc(&keys, cap * sizeof(int)));
        LPMB_CUDA(cudaMalloc(&strain, cap * sizeof(double)));
        brittle_candidates_kernel<<<g, DT, 0, c->stream>>>(N, Np, nn, nbi, fptr<double>(c, "dL"), fptr<double>(c, "distance_initial"),
                                                           param(c, "critical_bstrain"), cap, d_count, keys, strain, lpmb_own0(c), lpmb_own1(c));
        LPMB_LAUNCH_CHECK(c);
        int k = 0;
        LPMB_CUDA(cudaMemcpyAsync(&k, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        // this rank's candidates in scan order (i, then j ascending == key ascending)
        std::vector<long long> bi;
        std::vector<double> bs;
        bool overflow = k > cap;
        if (!overflow && k > 0) {
            std::vector<int> hk(k);
            std::vector<double> hs(k);
            LPMB_D2H(c, hk.data(), keys, k * sizeof(int));
            LPMB_D2H(c, hs.data(), strain, k * sizeof(double));
            std::vector<int> ord(k);
            for (int t = 0; t < k; t++)
                ord[t] = t;
            std::sort(ord.begin(), ord.end(), [&](int a, int b2) { return hk[a] < hk[b2]; });   // keys are unique
            bi.resize(k);
            bs.resize(k);
            for (int t = 0; t < k; t++) {
                bi[t] = hk[ord[t]];
                bs[t] = hs[ord[t]];
            }
        }
        long long gshift = 0;   // global index of a local particle = local index + gshift
        if (c->world > 1) {
            // sizes: {owned particles, candidates (-1: overflow)} of every rank
            LPMB_TRY(lpmb_ensure_staging(c, (size_t)(c->world + 1) * 2 * sizeof(long long)));
            long long mine[2] = {(long long)(lpmb_own1(c) - lpmb_own0(c)), overflow ? -1ll : (long long)k};
            long long *d_mine = (long long *)c->staging, *d_all = d_mine + 2;
            LPMB_H2D(c, d_mine, mine, sizeof(mine));
            LPMB_TRY(lpmb_dist_allgather_bytes(c, d_mine, d_all, sizeof(mine)));
            std::vector<long long> all((size_t)c->world * 2);
            LPMB_D2H(c, all.data(), d_all, all.size() * sizeof(long long));
            long long g0 = 0, kmax = 0, ktot = 0;
            for (int r = 0; r < c->world; r++) {
                if (r < c->rank)
                    g0 += all[2 * r];
                overflow = overflow || all[2 * r + 1] < 0;
                kmax = std::max(kmax, all[2 * r + 1]);
                ktot += std::max(0ll, all[2 * r + 1]);
            }
            overflow = overflow || ktot > cap;
            gshift = g0 - lpmb_own0(c);
            if (!overflow && ktot > 0) {
                // keys (global, 64-bit) and strains, padded to the longest list
                const size_t per = (size_t)kmax * 16;
                void *d_send = nullptr, *d_recv = nullptr;
                LPMB_CUDA(cudaMalloc(&d_send, per));
                LPMB_CUDA(cudaMalloc(&d_recv, per * c->world));
                std::vector<long long> sk(kmax, 0);
                std::vector<double> ss(kmax, 0.0);
                for (int t = 0; t < k; t++) {
                    sk[t] = (bi[t] / nn + gshift) * nn + bi[t] % nn;
                    ss[t] = bs[t];
                }
                LPMB_H2D(c, d_send, sk.data(), (size_t)kmax * 8);
                LPMB_H2D(c, (char *)d_send + (size_t)kmax * 8, ss.data(), (size_t)kmax * 8);
                int rc2 = lpmb_dist_allgather_bytes(c, d_send, d_recv, per);
                std::vector<char> h(per * c->world);
                if (rc2 == LPMB_OK)
                    LPMB_D2H(c, h.data(), d_recv, h.size());
                cudaFree(d_send);
                cudaFree(d_recv);
                if (rc2 != LPMB_OK) {
                    cudaFree(keys);
                    cudaFree(strain);
                    cudaFree(d_count);
                    return rc2;
                }
                bi.clear();
                bs.clear();
                for (int r = 0; r < c->world; r++) {
                    const long long *rk = (const long long *)(h.data() + per * r);
                    const double *rs = (const double *)(h.data() + per * r + (size_t)kmax * 8);
                    for (long long t = 0; t < all[2 * r + 1]; t++) {
                        bi.push_back(rk[t]);
                        bs.push_back(rs[t]);
                    }
                }
            }
            k = (int)ktot;
        }
        if (overflow) {
            cudaFree(keys);
            cudaFree(strain);
            cudaFree(d_count);
            lpmb_set_error("updateBrittleDamage: more than %d candidate bonds (the reference's b_cr[] holds 400, constitutive.c:1444)", cap);
            return LPMB_ERR_UNSUPPORTED;
        }
        *broken_out = k;  // the reference returns the candidate count even when it breaks only nbreak of them
        if (k > 0) {
            int first = 0;
            LPMB_TRY(lpmb_brittle_select(k, bi.data(), bs.data(), (int)param(c, "nbreak"), &first));
            const int nb = k - first;
            // winners on this rank: owned particles and the ghosts whose stars are complete (same slot order as on their owner)
            const int lo = c->world > 1 ? lpmb_own0(c) - c->narrow_recv_lo : 0, hi = c->world > 1 ? lpmb_own1(c) + c->narrow_recv_hi : N;
            std::vector<int> mine;
            for (int t = 0; t < nb; t++) {
                const long long li = bi[first + t] / nn - gshift;
                if (li >= lo && li < hi)
                    mine.push_back((int)li * nn + (int)(bi[first + t] % nn));
            }
            if (!mine.empty()) {
                LPMB_H2D(c, keys, mine.data(), mine.size() * sizeof(int));
                brittle_apply_kernel<<<lpmb_blocks((int)mine.size(), 128), 128, 0, c->stream>>>((int)mine.size(), nn, Np, keys, broken, dD0, w);
                c->launches++;
            }
            if (pairs && max_pairs > 0) {
                // (particle, neighbour) of every broken bond in GLOBAL indices; the neighbour is -1 where this rank does not hold the particle
                std::vector<int> hnbr((size_t)nn * Np);
                LPMB_D2H(c, hnbr.data(), nbr, hnbr.size() * sizeof(int));
                for (int t = 0; t < nb && t < max_pairs; t++) {
                    const long long gi = bi[first + t] / nn, li = gi - gshift;
                    const int j = (int)(bi[first + t] % nn);
                    pairs[2 * t] = (int)gi;
                    pairs[2 * t + 1] = (li >= 0 && li < N) ? (int)(hnbr[(size_t)j * Np + li] + gshift) : -1;
                }
            }
        }
        cudaFree(keys);
        cudaFree(strain);
    } else if (plmode == 5) {
        if (c->world != 1 || !c->params.count("damage_threshold") || !c->params.count("damagec_A")) {
            cudaFree(d_count);
            lpmb_set_error(c->world != 1 ? "updateDuctileDamageBwiseLocal is single-GPU only" : "damage_threshold / damagec_A not set");
            return c->world != 1 ? LPMB_ERR_UNSUPPORTED : LPMB_ERR_STATE;
        }
        const double thr = param(c, "damage_threshold");
        signed char *newly = nullptr;
        LPMB_CUDA(cudaMalloc(&newly, (size_t)nn * Np));
        LPMB_CUDA(cudaMemsetAsync(newly, 0, (size_t)nn * Np, c->stream));
        local_damage_kernel<<<g, DT, 0, c->stream>>>(N, thr, param(c, "damagec_A"), fptr<double>(c, "J2_triaxiality"), fptr<double>(c, "J2_dlambda"),
                                                     fptr<double>(c, "damage_local0"));
        LPMB_LAUNCH_CHECK(c);
        local_break_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, thr, fptr<double>(c, "damage_local0"), broken, dD0, fptr<int>(c, "nb"), newly,
                                                    d_count);
        LPMB_LAUNCH_CHECK(c);
        local_finish_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, fptr<int>(c, "nb"), broken, dD0, w);
        LPMB_LAUNCH_CHECK(c);
        int k = 0;
        LPMB_CUDA(cudaMemcpyAsync(&k, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        *broken_out = k;
        if (k > 0 && pairs && max_pairs > 0) {
            // reference logging order: i ascending, then slot j ascending (constitutive.c:1644-1676)
            std::vector<signed char> hn((size_t)nn * Np);
            std::vector<int> hnbr((size_t)nn * Np);
            LPMB_D2H(c, hn.data(), newly, hn.size());
            LPMB_D2H(c, hnbr.data(), nbr, hnbr.size() * sizeof(int));
            int t = 0;
            for (int i = 0; i < N && t < max_pairs; i++)
                for (int j = 0; j < nn && t < max_pairs; j++)
                    if (hn[(size_t)j * Np + i]) {
                        pairs[2 * t] = i;
                        pairs[2 * t + 1] = hnbr[(size_t)j * Np + i];
                        t++;
                    }
        }
        cudaFree(newly);
    }
    else if (plmode == LPMB_DAMAGE_PWISE_LOCAL || plmode == LPMB_DAMAGE_BWISE_NONLOCAL) {
        const bool pw = plmode == LPMB_DAMAGE_PWISE_LOCAL;
        const bool have = c->params.count("damage_threshold") && c->params.count("damagec_A") &&
                          (pw || (c->params.count("damage_L") && c->params.count("particle_volume")));
        const signed char *mirror = fptr<signed char>(c, "mirror");
        const double *L0 = fptr<double>(c, "distance_initial");
        if (c->world != 1 || !have || (pw ? !mirror : !L0)) {
            cudaFree(d_count);
            lpmb_set_error(c->world != 1 ? "this damage law is single-GPU only" : "damage parameters / topology fields not set");
            return c->world != 1 ? LPMB_ERR_UNSUPPORTED : LPMB_ERR_STATE;
        }
        const double thr = param(c, "damage_threshold"), Ac = param(c, "damagec_A");
        signed char *newly = nullptr;  // pw: one flag per particle; else one per bond slot
        const size_t nflag = pw ? (size_t)Np : (size_t)nn * Np;
        LPMB_CUDA(cudaMalloc(&newly, nflag));
        LPMB_CUDA(cudaMemsetAsync(newly, 0, nflag, c->stream));
        if (pw) {
            pwise_local_damage_kernel<<<g, DT, 0, c->stream>>>(N, thr, Ac, fptr<double>(c, "J2_triaxiality"), fptr<double>(c, "J2_dlambda"),
                                                               fptr<double>(c, "damage_local0"), newly, d_count);
            LPMB_LAUNCH_CHECK(c);
            pwise_local_break_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, mirror, fptr<double>(c, "damage_local0"), newly, broken, dD0, w);
            LPMB_LAUNCH_CHECK(c);
        } else {
            bwise_nonlocal_damage_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, L0, param(c, "damage_L"), thr, Ac, param(c, "particle_volume"),
                                                                  fptr<double>(c, "J2_dlambda"), fptr<double>(c, "J2_triaxiality"),
                                                                  fptr<double>(c, "damage_nonlocal0"));
            LPMB_LAUNCH_CHECK(c);
            bwise_nonlocal_break_kernel<<<g, DT, 0, c->stream>>>(N, Np, nbi, nbr, thr, fptr<double>(c, "damage_nonlocal0"), broken, dD0, w, newly,
                                                                 d_count);
            LPMB_LAUNCH_CHECK(c);
        }
        int k = 0;
        LPMB_CUDA(cudaMemcpyAsync(&k, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        *broken_out = k;
        if (k > 0 && pairs && max_pairs > 0) {
            // reference logging order: particle (or i, then slot j) ascending; the particle-wise law logs a single
            // index per entry (constitutive.c:1562) -> second column -1
            std::vector<signed char> hn(nflag);
            LPMB_D2H(c, hn.data(), newly, hn.size());
            int t = 0;
            if (pw) {
                for (int i = 0; i < N && t < max_pairs; i++)
                    if (hn[i]) {
                        pairs[2 * t] = i;
                        pairs[2 * t + 1] = -1;
                        t++;
                    }
            } else {
                std::vector<int> hnbr((size_t)nn * Np);
                LPMB_D2H(c, hnbr.data(), nbr, hnbr.size() * sizeof(int));
                for (int i = 0; i < N && t < max_pairs; i++)
                    for (int j = 0; j < nn && t < max_pairs; j++)
                        if (hn[(size_t)j * Np + i]) {
                            pairs[2 * t] = i;
                            pairs[2 * t + 1] = hnbr[(size_t)j * Np + i];
                            t++;
                        }
            }
        }
        cudaFree(newly);
    }
    // any other plmode: the reference's dispatcher does nothing and returns 0 (constitutive.c:149-164)
    cudaFree(d_count);
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}
