// lpmb_mg.cu -- matrix-free geometric multigrid V-cycle: the preconditioner of the OPT-IN fast mode of the solve
// (param "cg_precond" = 1; lpmb_solver.cu::pcg_run).  The parity mode -- the reference's unpreconditioned CG,
// solver.c:188-270 with ipar[10] = 0 -- is untouched by anything in this file.
//
// Why.  north_star asks for a preconditioned CG.  On the lattice tangent point preconditioners do not pay (Jacobi /
// block-Jacobi save 4-5 % of the iterations, tests/test_oracle_ref.py); the iteration count grows like n^(1/3) (116 at
// 48^3, 226 at 100^3, 458 at 216^3), i.e. it is a multilevel problem.  What makes a multilevel preconditioner nearly
// free here: the reference assembles K as the ELASTIC tangent only (stiffness.c:394,420: plmode 6), so on an undamaged
// simple-cubic block every interior block row of K is the same 61-point stencil of 3x3 blocks up to the O(strain)
// geometric change -- the hierarchy needs NO matrices, one stencil in constant memory serves all levels:
//
//   level l = the simple-cubic lattice coarsened l times (vertex-centred: coarse site I sits on fine site 2I),
//   A_l u (i)  =  sum over stencil offsets o with i+o inside the box of  2^l S_o (u(i+o) - u(i))
//                 -- S_o = the off-diagonal blocks of one interior row of the assembled K; the diagonal block is minus
//                 the sum of the off-diagonal blocks that are present, which is what translation invariance of the real
//                 tangent gives (free surfaces handled); the factor 2^l is the rediscretised operator on spacing 2^l h
//                 (bond stiffness ~ radius, stiffness.c:179-185), equal to the Galerkin operator on smooth fields,
//   smoother   =  damped block-Jacobi (3x3 diagonal blocks inverted per boundary class on the host), nu = 2 sweeps before
//                 and after the coarse correction with the two damping factors of the degree-2 Chebyshev polynomial
//                 (0.56, 1.39), R = P^T with trilinear P: a symmetric positive definite V-cycle,
//   constraints: the DoF mask of the solve on level 0; a coarse DoF is constrained when any fine DoF in its
//                 interpolation support is (Dirichlet faces stay Dirichlet faces).
//
// numpy prototype on the reference's own tangent (C5 material, 1 % stretch, bottom layer held): stencil operator vs real K
// 0.4-0.6 % apart; PCG iterations to the reference's stop rule 11 at 24^3 (plain CG 64), 13 at 48^3 (plain CG 116).
//
// Cost per V-cycle at 216^3: 5 fine-level stencil passes over 30 M unknowns with no matrix traffic (x is read through
// L1/L2, 61 x 18 flop per site) + 1/7 of that for the coarse levels -- against 24 GB of matrix per real SpMV.
//
// Scope: 3-D simple-cubic FULL blocks numbered x-fastest (initialization.c:266-284) with >= 5 sites per edge, on one GPU or
// on z-slabs (below); anything else -> LPMB_ERR_UNSUPPORTED (the caller fails loudly; no silent fallback to another solver).
//
// Slab runs (world > 1, param mg_dist = 1, default): ONE global hierarchy, distributed.  Coarse site I sits on fine site 2I
// in GLOBAL lattice coordinates, a rank owns the coarse layers whose fine layer it owns, and every distributed level keeps
// [2 ghost layers | owned layers | 2 ghost layers] (level 0: the context's own 4-layer ghosts).  Before every stencil pass
// (smoother sweep, residual) the two ghost layers of the iterate are refreshed from the neighbours, the residual's before
// the restriction, the coarse correction's before the prolongation -- 5 small exchanges per level and V-cycle; boundary
// classes and interpolation weights use global coordinates, so the operator is the single-GPU one.  From the first level on
// which some rank would own fewer than 2 layers the level is REPLICATED: the ranks' restricted residuals are all-gathered and
// every rank runs the remaining (tiny) levels redundantly.  All ranks use the stencil read by the middle rank (a symmetric
// preconditioner needs one operator).  mg_dist = 0 keeps a hierarchy per slab without any communication: block-Jacobi over
// the slabs, symmetric positive definite but without coupling across the interfaces (41 instead of 9 iterations on two slabs
// of the 216^3 block, profiles/r02af_bench_n2.log).
#include <algorithm>
#include <cmath>

#include "lpmb_internal.cuh"

#define MG_MAXOFF 64
#define MG_MAXLEV 12

__constant__ int c_mg_off[MG_MAXOFF][3];
__constant__ double c_mg_S[MG_MAXOFF][9];
__constant__ int c_mg_noff;
__constant__ int c_mg_delta[MG_MAXOFF];   // shared-memory slot offset of each stencil offset inside the 36 x 8 x 8 tile

struct MGLevel {
    int nx = 0, ny = 0, nz = 0;
    long long n = 0, stride = 0;   // sites; distance between the components of a vector
    double scale = 1.0;
    double *dinv = nullptr;        // [729][9] inverse diagonal block per boundary class
    double *mask = nullptr;        // [3][stride] (level 0: the solve's mask, not owned)
    double *u = nullptr, *u2 = nullptr, *f = nullptr, *res = nullptr;   // [3][stride] (level 0: u = caller's z, f = caller's r)
    // slab runs: nz above = layers of the LOCAL block of this level
    int gz0 = 0;            // global z index of local layer 0
    int nzg = 0;            // layers of the whole level (== nz on one GPU and on replicated levels)
    int oz0 = 0, oz1 = 0;   // owned local layers [oz0, oz1)
    bool dist = false;      // distributed over the ranks (ghost layers refreshed by exchanges)
};

struct MGState {
    bool ready = false;
    int nlev = 0;
    MGLevel lev[MG_MAXLEV];
    double S[MG_MAXOFF][9];
    int off[MG_MAXOFF][3];
    int noff = 0;
    int delta[MG_MAXOFF];
    int nu = 2, nu_coarse = 40;
    double omega = 0.56, omega2 = 1.39;   // damping of the odd / even sweeps (two different values = a degree-2 polynomial smoother)
    // slab runs
    bool dist = false;                    // distributed hierarchy (mg_dist = 1)
    double *mask_phys = nullptr;          // [3][Np] level-0 mask of the boundary conditions alone (the solve's mask also zeroes the ghosts)
    int lrep = -1;                        // first replicated level
    long long gat_off[LPMB_PEER_MAXW], gat_cnt[LPMB_PEER_MAXW];   // the ranks' owned ranges of level lrep (elements per component)
};

static std::map<lpmb_ctx *, MGState> g_mg;
static lpmb_ctx *g_mg_const_owner[64] = {nullptr};   // per device: the context whose stencil sits in __constant__ memory

void lpmb_mg_release(lpmb_ctx *c)
{
    auto it = g_mg.find(c);
    if (it == g_mg.end())
        return;
    MGState &M = it->second;
    if (c->device >= 0 && c->device < 64 && g_mg_const_owner[c->device] == c)
        g_mg_const_owner[c->device] = nullptr;
    cudaFree(M.mask_phys);
    for (int l = 0; l < M.nlev; l++) {
        MGLevel &L = M.lev[l];
        cudaFree(L.dinv);
        cudaFree(L.u2);
        cudaFree(L.res);
        if (l > 0) {
            cudaFree(L.mask);
            cudaFree(L.u);
            cudaFree(L.f);
        }
    }
    g_mg.erase(it);
}

void lpmb_mg_touch(lpmb_ctx *c)   // K.val changed: the stencil is re-read at the next solve
{
    auto it = g_mg.find(c);
    if (it != g_mg.end())
        it->second.ready = false;
}

// ---- kernels ---------------------------------------------------------------------------------------
// a += S_k d.  (Skipping the zero entries of the mostly rank-1 stencil blocks with warp-uniform branches was measured and
// is SLOWER -- 415 instead of 335 us per tiled pass on average, profiles/r02l_fast_launches.csv: the dense form keeps the
// constant-bank operands fused into nine back-to-back DFMAs.)
#define MG_ACC(k, d0, d1, d2, a0, a1, a2)                                                        \
    do {                                                                                         \
        a0 = fma(c_mg_S[k][0], d0, fma(c_mg_S[k][1], d1, fma(c_mg_S[k][2], d2, a0)));            \
        a1 = fma(c_mg_S[k][3], d0, fma(c_mg_S[k][4], d1, fma(c_mg_S[k][5], d2, a1)));            \
        a2 = fma(c_mg_S[k][6], d0, fma(c_mg_S[k][7], d1, fma(c_mg_S[k][8], d2, a2)));            \
    } while (0)

__device__ __forceinline__ int mg_axis_class(int i, int n) { return min(i, 2) + 3 * min(n - 1 - i, 2); }

// MODE 0: out = u + omega * mask .* Dinv (f - A u)      (one damped block-Jacobi sweep, out != u)
// MODE 1: out = mask .* (f - A u)                       (residual)
// MODE 2: out = omega * mask .* Dinv f                  (first sweep from u = 0: no stencil pass)
template <int MODE>
__global__ void __launch_bounds__(128)
mg_stencil_kernel(int nx, int ny, int nz, int gz0, int nzg, long long stride, double scale, const double *__restrict__ dinv,
                  const double *__restrict__ mask, const double *__restrict__ u, const double *__restrict__ f, double *__restrict__ out, double omega,
                  const double *__restrict__ done)
{
    if (done && done[0] != 0.0)
        return;
    const long long n = (long long)nx * ny * nz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int ix = (int)(i % nx), iy = (int)((i / nx) % ny), iz = (int)(i / ((long long)nx * ny));
    double r0 = f[i], r1 = f[stride + i], r2 = f[2 * stride + i];
    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
    if (MODE != 2) {
        u0 = u[i], u1 = u[stride + i], u2 = u[2 * stride + i];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const int noff = c_mg_noff;
        for (int k = 0; k < noff; k++) {
            const int jx = ix + c_mg_off[k][0], jy = iy + c_mg_off[k][1], jz = iz + c_mg_off[k][2];
            if (jx < 0 || jx >= nx || jy < 0 || jy >= ny || jz < 0 || jz >= nz)
                continue;
            const long long j = jx + (long long)nx * (jy + (long long)ny * jz);
            const double d0 = __ldg(u + j) - u0, d1 = __ldg(u + stride + j) - u1, d2 = __ldg(u + 2 * stride + j) - u2;
            MG_ACC(k, d0, d1, d2, a0, a1, a2);
        }
        r0 -= scale * a0;
        r1 -= scale * a1;
        r2 -= scale * a2;
    }
    const double m0 = mask ? mask[i] : 1.0, m1 = mask ? mask[stride + i] : 1.0, m2 = mask ? mask[2 * stride + i] : 1.0;
    if (MODE == 1) {
        out[i] = m0 * r0;
        out[stride + i] = m1 * r1;
        out[2 * stride + i] = m2 * r2;
        return;
    }
    // constrained DoFs carry no residual: the block solve must not leak their (unmasked) residual into the free ones
    r0 *= m0, r1 *= m1, r2 *= m2;
    const double *D = dinv + 9 * (mg_axis_class(ix, nx) + 9 * (mg_axis_class(iy, ny) + 9 * mg_axis_class(iz + gz0, nzg)));   // the class is a property of the GLOBAL position
    out[i] = u0 + omega * m0 * (D[0] * r0 + D[1] * r1 + D[2] * r2);
    out[stride + i] = u1 + omega * m1 * (D[3] * r0 + D[4] * r1 + D[5] * r2);
    out[2 * stride + i] = u2 + omega * m2 * (D[6] * r0 + D[7] * r1 + D[8] * r2);
}

// The same three operations with the u values of a 32 x 4 x 4 tile of sites + its 2-site apron staged in shared memory
// (36 x 8 x 8 sites x 3 components = 55 KB): a site is read from L2 4.5 times per pass instead of 61 times.  Used on
// the levels large enough for the staging to pay.  Out-of-box apron sites hold 0 and are skipped by the same coordinate
// test as in the plain kernel; blocks whose whole apron lies inside the box skip the tests.
#define MG_TX 32
#define MG_TY 4
#define MG_TZ 4
#define MG_SX (MG_TX + 4)
#define MG_SY (MG_TY + 4)
#define MG_SZ (MG_TZ + 4)
#define MG_TILE_SITES (MG_SX * MG_SY * MG_SZ)
#define MG_SC_NOFF 60   // off-diagonal blocks of an interior simple-cubic block row (61 conn entries minus the particle itself)
template <int MODE>
__global__ void __launch_bounds__(MG_TX * MG_TY * MG_TZ)
mg_stencil_tiled_kernel(int nx, int ny, int nz, int gz0, int nzg, long long stride, double scale, const double *__restrict__ dinv,
                        const double *__restrict__ mask, const double *__restrict__ u, const double *__restrict__ f, double *__restrict__ out,
                        double omega, const double *__restrict__ done)
{
    extern __shared__ double mg_us[];   // [3][MG_TILE_SITES]
    if (done && done[0] != 0.0)
        return;
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int x0 = blockIdx.x * MG_TX, y0 = blockIdx.y * MG_TY, z0 = blockIdx.z * MG_TZ;
    const int tid = tx + MG_TX * (ty + MG_TY * tz);
    for (int s = tid; s < MG_TILE_SITES; s += MG_TX * MG_TY * MG_TZ) {
        const int sx = s % MG_SX, sy = (s / MG_SX) % MG_SY, sz = s / (MG_SX * MG_SY);
        const int gx = x0 + sx - 2, gy = y0 + sy - 2, gz = z0 + sz - 2;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz) {
            const long long j = gx + (long long)nx * (gy + (long long)ny * gz);
            v0 = u[j], v1 = u[stride + j], v2 = u[2 * stride + j];
        }
        mg_us[s] = v0;
        mg_us[MG_TILE_SITES + s] = v1;
        mg_us[2 * MG_TILE_SITES + s] = v2;
    }
    __syncthreads();
    const int ix = x0 + tx, iy = y0 + ty, iz = z0 + tz;
    if (ix >= nx || iy >= ny || iz >= nz)
        return;
    const long long i = ix + (long long)nx * (iy + (long long)ny * iz);
    const int me = (tx + 2) + MG_SX * ((ty + 2) + MG_SY * (tz + 2));
    const double u0 = mg_us[me], u1 = mg_us[MG_TILE_SITES + me], u2 = mg_us[2 * MG_TILE_SITES + me];
    const bool interior = x0 >= 2 && x0 + MG_TX + 2 <= nx && y0 >= 2 && y0 + MG_TY + 2 <= ny && z0 >= 2 && z0 + MG_TZ + 2 <= nz;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int noff = c_mg_noff;
    if (interior && noff == MG_SC_NOFF) {
        // the common case (simple-cubic 2-hop stencil, tile away from the faces): fully unrolled, slot offsets and stencil
        // entries are constant-bank operands of the instructions -- no index arithmetic, no bounds tests
#pragma unroll
        for (int k = 0; k < MG_SC_NOFF; k++) {
            const int sl = me + c_mg_delta[k];
            const double d0 = mg_us[sl] - u0, d1 = mg_us[MG_TILE_SITES + sl] - u1, d2 = mg_us[2 * MG_TILE_SITES + sl] - u2;
            MG_ACC(k, d0, d1, d2, a0, a1, a2);
        }
    } else {
        for (int k = 0; k < noff; k++) {
            const int ox = c_mg_off[k][0], oy = c_mg_off[k][1], oz = c_mg_off[k][2];
            if (!interior) {
                const int jx = ix + ox, jy = iy + oy, jz = iz + oz;
                if (jx < 0 || jx >= nx || jy < 0 || jy >= ny || jz < 0 || jz >= nz)
                    continue;
            }
            const int sl = me + ox + MG_SX * (oy + MG_SY * oz);
            const double d0 = mg_us[sl] - u0, d1 = mg_us[MG_TILE_SITES + sl] - u1, d2 = mg_us[2 * MG_TILE_SITES + sl] - u2;
            MG_ACC(k, d0, d1, d2, a0, a1, a2);
        }
    }
    double r0 = f[i] - scale * a0, r1 = f[stride + i] - scale * a1, r2 = f[2 * stride + i] - scale * a2;
    const double m0 = mask ? mask[i] : 1.0, m1 = mask ? mask[stride + i] : 1.0, m2 = mask ? mask[2 * stride + i] : 1.0;
    if (MODE == 1) {
        out[i] = m0 * r0;
        out[stride + i] = m1 * r1;
        out[2 * stride + i] = m2 * r2;
        return;
    }
    r0 *= m0, r1 *= m1, r2 *= m2;
    const double *D = dinv + 9 * (mg_axis_class(ix, nx) + 9 * (mg_axis_class(iy, ny) + 9 * mg_axis_class(iz + gz0, nzg)));   // the class is a property of the GLOBAL position
    out[i] = u0 + omega * m0 * (D[0] * r0 + D[1] * r1 + D[2] * r2);
    out[stride + i] = u1 + omega * m1 * (D[3] * r0 + D[4] * r1 + D[5] * r2);
    out[2 * stride + i] = u2 + omega * m2 * (D[6] * r0 + D[7] * r1 + D[8] * r2);
}

template <int MODE>
static int mg_launch_stencil(lpmb_ctx *c, const MGLevel &L, const double *u, const double *f, double *out, double omega, const double *done)
{
    const bool tiled = MODE != 2 && (double)L.n >= param(c, "mg_tiled_min", 1024.0) && param(c, "mg_tiled", 1.0) != 0.0;
    if (tiled) {
        const size_t smem = (size_t)3 * MG_TILE_SITES * sizeof(double);
        // > 48 KB of dynamic shared memory: opt in (per device; cheap enough to repeat)
        LPMB_CUDA(cudaFuncSetAttribute(mg_stencil_tiled_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((L.nx + MG_TX - 1) / MG_TX, (L.ny + MG_TY - 1) / MG_TY, (L.nz + MG_TZ - 1) / MG_TZ), block(MG_TX, MG_TY, MG_TZ);
        mg_stencil_tiled_kernel<MODE><<<grid, block, smem, c->stream>>>(L.nx, L.ny, L.nz, L.gz0, L.nzg, L.stride, L.scale, L.dinv, L.mask, u, f, out, omega, done);
    } else {
        mg_stencil_kernel<MODE><<<lpmb_blocks(L.n, 128), 128, 0, c->stream>>>(L.nx, L.ny, L.nz, L.gz0, L.nzg, L.stride, L.scale, L.dinv, L.mask, u, f, out, omega, done);
    }
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

// Coarsest level (<= MG_COARSE_MAX sites): all nu damped-Jacobi sweeps from u = 0 in ONE block, u ping-ponging in shared
// memory -- the same arithmetic as nu launches of the plain kernel (MODE 2, then MODE 0), without 40 launch latencies.
#define MG_COARSE_MAX 512
__global__ void __launch_bounds__(256)
mg_coarse_solve_kernel(int nx, int ny, int nz, long long stride, double scale, const double *__restrict__ dinv, const double *__restrict__ mask,
                       const double *__restrict__ f, double *__restrict__ out, double omega, int nu, const double *__restrict__ done)
{
    __shared__ double us[2][3][MG_COARSE_MAX];
    if (done && done[0] != 0.0)
        return;
    const int n = nx * ny * nz;
    for (int sweep = 0; sweep < nu; sweep++) {
        const int src = (sweep & 1) ^ 1, dst = sweep & 1;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
            double r0 = f[i], r1 = f[stride + i], r2 = f[2 * stride + i];
            double u0 = 0.0, u1 = 0.0, u2 = 0.0;
            if (sweep > 0) {
                u0 = us[src][0][i], u1 = us[src][1][i], u2 = us[src][2][i];
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                const int noff = c_mg_noff;
                for (int k = 0; k < noff; k++) {
                    const int jx = ix + c_mg_off[k][0], jy = iy + c_mg_off[k][1], jz = iz + c_mg_off[k][2];
                    if (jx < 0 || jx >= nx || jy < 0 || jy >= ny || jz < 0 || jz >= nz)
                        continue;
                    const int j = jx + nx * (jy + ny * jz);
                    const double d0 = us[src][0][j] - u0, d1 = us[src][1][j] - u1, d2 = us[src][2][j] - u2;
                    MG_ACC(k, d0, d1, d2, a0, a1, a2);
                }
                r0 -= scale * a0;
                r1 -= scale * a1;
                r2 -= scale * a2;
            }
            const double m0 = mask[i], m1 = mask[stride + i], m2 = mask[2 * stride + i];
            r0 *= m0, r1 *= m1, r2 *= m2;
            const double *D = dinv + 9 * (mg_axis_class(ix, nx) + 9 * (mg_axis_class(iy, ny) + 9 * mg_axis_class(iz, nz)));
            us[dst][0][i] = u0 + omega * m0 * (D[0] * r0 + D[1] * r1 + D[2] * r2);
            us[dst][1][i] = u1 + omega * m1 * (D[3] * r0 + D[4] * r1 + D[5] * r2);
            us[dst][2][i] = u2 + omega * m2 * (D[6] * r0 + D[7] * r1 + D[8] * r2);
        }
        __syncthreads();
    }
    const int last = (nu - 1) & 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        out[i] = us[last][0][i];
        out[stride + i] = us[last][1][i];
        out[2 * stride + i] = us[last][2][i];
    }
}

// 1-D interpolation weight of fine site f from coarse site X (coarse X sits on fine 2X; nc coarse sites)
__device__ __forceinline__ double mg_w1(int f, int X, int nc)
{
    const int d = f - 2 * X;
    if (d == 0)
        return 1.0;
    if (d == 1)
        return X + 1 < nc ? 0.5 : 1.0;   // the upper coarse neighbour does not exist: constant extrapolation
    if (d == -1)
        return 0.5;
    return 0.0;
}

// fc = mask_c .* P^T rf     (full weighting = transpose of the trilinear interpolation)
__global__ void __launch_bounds__(128)
mg_restrict_kernel(int nfx, int nfy, int nfz, long long sf, int ncx, int ncy, int ncz, long long sc, int gzf, int gzc, int nczg,
                   const double *__restrict__ rf, const double *__restrict__ mask_c, double *__restrict__ fc, const double *__restrict__ done)
{
    if (done && done[0] != 0.0)
        return;
    const long long nc = (long long)ncx * ncy * ncz;
    const long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc)
        return;
    const int X = (int)(I % ncx), Y = (int)((I / ncx) % ncy), Z = (int)(I / ((long long)ncx * ncy));
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const int Zg = Z + gzc;   // global coarse layer; its fine layer 2 Zg is local fine layer 2 Zg - gzf
    for (int dz = -1; dz <= 1; dz++) {
        const int fz = 2 * Zg + dz - gzf;
        if (fz < 0 || fz >= nfz)
            continue;
        const double wz = mg_w1(fz + gzf, Zg, nczg);
        for (int dy = -1; dy <= 1; dy++) {
            const int fy = 2 * Y + dy;
            if (fy < 0 || fy >= nfy)
                continue;
            const double wy = wz * mg_w1(fy, Y, ncy);
            for (int dx = -1; dx <= 1; dx++) {
                const int fx = 2 * X + dx;
                if (fx < 0 || fx >= nfx)
                    continue;
                const double w = wy * mg_w1(fx, X, ncx);
                const long long j = fx + (long long)nfx * (fy + (long long)nfy * fz);
                s0 = fma(w, rf[j], s0);
                s1 = fma(w, rf[sf + j], s1);
                s2 = fma(w, rf[2 * sf + j], s2);
            }
        }
    }
    fc[I] = mask_c[I] * s0;
    fc[sc + I] = mask_c[sc + I] * s1;
    fc[2 * sc + I] = mask_c[2 * sc + I] * s2;
}

// uf += mask_f .* P uc
__global__ void __launch_bounds__(128)
mg_prolong_kernel(int nfx, int nfy, int nfz, long long sf, int ncx, int ncy, int ncz, long long sc, int gzf, int gzc, int nczg,
                  const double *__restrict__ uc, const double *__restrict__ mask_f, double *__restrict__ uf, const double *__restrict__ done)
{
    if (done && done[0] != 0.0)
        return;
    const long long nf = (long long)nfx * nfy * nfz;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf)
        return;
    const int fx = (int)(i % nfx), fy = (int)((i / nfx) % nfy), fz = (int)(i / ((long long)nfx * nfy));
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const int fzg = fz + gzf;   // global fine layer
    const int X0 = fx >> 1, Y0 = fy >> 1, Z0 = fzg >> 1;
    for (int az = 0; az <= (fzg & 1); az++) {
        const int Zg = Z0 + az, Z = Zg - gzc;
        if (Zg >= nczg || Z < 0 || Z >= ncz)
            continue;
        const double wz = mg_w1(fzg, Zg, nczg);
        for (int ay = 0; ay <= (fy & 1); ay++) {
            const int Y = Y0 + ay;
            if (Y >= ncy)
                continue;
            const double wy = wz * mg_w1(fy, Y, ncy);
            for (int ax = 0; ax <= (fx & 1); ax++) {
                const int X = X0 + ax;
                if (X >= ncx)
                    continue;
                const double w = wy * mg_w1(fx, X, ncx);
                const long long J = X + (long long)ncx * (Y + (long long)ncy * Z);
                s0 = fma(w, uc[J], s0);
                s1 = fma(w, uc[sc + J], s1);
                s2 = fma(w, uc[2 * sc + J], s2);
            }
        }
    }
    const double m0 = mask_f ? mask_f[i] : 1.0, m1 = mask_f ? mask_f[sf + i] : 1.0, m2 = mask_f ? mask_f[2 * sf + i] : 1.0;
    uf[i] += m0 * s0;
    uf[sf + i] += m1 * s1;
    uf[2 * sf + i] += m2 * s2;
}

// coarse DoF free only if every fine DoF in its interpolation support is free
__global__ void mg_coarse_mask_kernel(int nfx, int nfy, int nfz, long long sf, int ncx, int ncy, int ncz, long long sc, int gzf, int gzc, int nczg,
                                      const double *__restrict__ mf, double *__restrict__ mc)
{
    const long long nc = (long long)ncx * ncy * ncz;
    const long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc)
        return;
    const int X = (int)(I % ncx), Y = (int)((I / ncx) % ncy), Z = (int)(I / ((long long)ncx * ncy));
    double m[3] = {1.0, 1.0, 1.0};
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int fx = 2 * X + dx, fy = 2 * Y + dy, fz = 2 * (Z + gzc) + dz - gzf;
                if (fx < 0 || fx >= nfx || fy < 0 || fy >= nfy || fz < 0 || fz >= nfz)
                    continue;
                if (mg_w1(fx, X, ncx) * mg_w1(fy, Y, ncy) * mg_w1(fz + gzf, Z + gzc, nczg) == 0.0)
                    continue;
                const long long j = fx + (long long)nfx * (fy + (long long)nfy * fz);
                for (int k = 0; k < 3; k++)
                    m[k] = fmin(m[k], mf ? mf[k * sf + j] : 1.0);
            }
    for (int k = 0; k < 3; k++)
        mc[k * sc + I] = m[k];
}

// particle i must sit on lattice site (i % nx, (i / nx) % ny, i / (nx ny)) of an axis-aligned lattice of spacing q
__global__ void mg_check_order_kernel(int N, int Np, const double *__restrict__ x0, double ox, double oy, double oz, double q, int nx, int ny,
                                      int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const double fx = (x0[i] - ox) / q, fy = (x0[(size_t)Np + i] - oy) / q, fz = (x0[(size_t)2 * Np + i] - oz) / q;
    const int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
    if (fabs(fx - ix) > 1e-6 || fabs(fy - iy) > 1e-6 || fabs(fz - iz) > 1e-6)
        atomicExch(bad, 1);
}

// ---- host side ---------------------------------------------------------------------------------------
static bool invert3(const double *a, double *inv)
{
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (!(std::fabs(det) > 0.0))
        return false;
    const double id = 1.0 / det;
    inv[0] = (a[4] * a[8] - a[5] * a[7]) * id;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) * id;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) * id;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return true;
}

// The multigrid levels of a full simple-cubic block -- and, for z-slabs, the part of every level a rank holds.  Pure host
// arithmetic (no device, no context): tests/test_partition.py checks its invariants for all ranks on CPU.
//   owned[r]   fine lattice layers rank r owns (world = 1: the whole block); nz_local0 / ghost_lo0: layers of this rank's
//              level-0 block and how many of them lie below its owned ones
//   plan[8 l + {0: distributed, 1: nx, 2: ny, 3: nz of the local block, 4: global z of local layer 0, 5 / 6: owned local
//              layers [oz0, oz1), 7: layers of the whole level}]
//   Coarse site I sits on fine site 2 I (global), so a rank owns the coarse layers [ceil(a / 2), ceil(b / 2)) of its fine
//   layers [a, b); a distributed level keeps 2 ghost layers either side (clipped at the block's faces).  The level is
//   replicated (lrep = its index; gat_off / gat_cnt = the ranks' owned ranges of it in elements per component) as soon as a
//   rank would own fewer than 2 layers or the level has <= 8 layers; levels end when an edge is <= 4 sites.
extern "C" int lpmb_mg_slab_plan(int world, int rank, const long long *owned, int nx, int ny, int nz_local0, int ghost_lo0, int max_levels, int *plan,
                                 long long *gat_off, long long *gat_cnt, int *nlev, int *lrep)
{
    LPMB_REQUIRE(world >= 1 && world <= LPMB_PEER_MAXW && rank >= 0 && rank < world && owned && plan && nlev && lrep && nx > 0 && ny > 0 &&
                     max_levels > 0 && (world == 1 || (gat_off && gat_cnt)),
                 LPMB_ERR_ARG, "lpmb_mg_slab_plan: bad argument");
    std::vector<long long> ga(world + 1, 0);
    for (int r = 0; r < world; r++) {
        LPMB_REQUIRE(owned[r] > 0, LPMB_ERR_ARG, "lpmb_mg_slab_plan: rank %d owns no layer", r);
        ga[r + 1] = ga[r] + owned[r];
    }
    int ax = nx, ay = ny, az = (int)ga[world], n = 0;
    bool level_dist = world > 1;
    *lrep = -1;
    for (;;) {
        LPMB_REQUIRE(n < max_levels, LPMB_ERR_UNSUPPORTED, "cg_precond: too many levels");
        int *P = plan + 8 * n;
        P[0] = level_dist ? 1 : 0, P[1] = ax, P[2] = ay, P[7] = az;
        const int A = (int)ga[rank], B = (int)ga[rank + 1];
        if (n == 0 && world > 1) {
            P[3] = nz_local0, P[4] = A - ghost_lo0, P[5] = ghost_lo0, P[6] = ghost_lo0 + (B - A);
        } else if (level_dist) {
            P[4] = std::max(0, A - 2);
            P[3] = std::min(az, B + 2) - P[4];
            P[5] = A - P[4], P[6] = B - P[4];
        } else {
            P[3] = az, P[4] = 0, P[5] = world > 1 ? A : 0, P[6] = world > 1 ? B : az;
        }
        n++;
        if (std::min(ax, std::min(ay, az)) <= 4 && !level_dist)
            break;
        LPMB_REQUIRE(std::min(ax, std::min(ay, az)) > 2, LPMB_ERR_UNSUPPORTED, "cg_precond on slabs: the block is too thin for %d ranks", world);
        ax = (ax + 1) / 2, ay = (ay + 1) / 2, az = (az + 1) / 2;
        for (int r = 0; r <= world; r++)
            ga[r] = (ga[r] + 1) / 2;
        if (level_dist) {
            long long min_own = 1ll << 40;
            for (int r = 0; r < world; r++)
                min_own = std::min(min_own, ga[r + 1] - ga[r]);
            if (min_own < 2 || az <= 8) {
                level_dist = false;
                *lrep = n;
                for (int r = 0; r < world; r++) {
                    gat_off[r] = ga[r] * (long long)ax * ay;
                    gat_cnt[r] = (ga[r + 1] - ga[r]) * (long long)ax * ay;
                }
            }
        }
    }
    *nlev = n;
    return LPMB_OK;
}

// lattice dimensions + ordering check (once), level storage
static int mg_build_levels(lpmb_ctx *c, MGState &M)
{
    // Slab runs: the hierarchy is built on the rank's LOCAL block [ghost layers | owned layers | ghost layers] -- itself a full
    // x-fastest simple-cubic block.  The level-0 mask constrains every ghost DoF, so the V-cycle is a multigrid solve of the
    // slab's own Dirichlet problem and the preconditioner of the whole system is block-Jacobi over the slabs (additive Schwarz
    // without a coarse space across ranks): symmetric positive definite, no communication inside the V-cycle.
    LPMB_REQUIRE(c->dim == 3 && c->lattice == LPMB_LATTICE_SC, LPMB_ERR_UNSUPPORTED, "cg_precond: 3-D simple-cubic lattices only");
    LPMB_REQUIRE(c->params.count("radius") && c->fields.count("xyz_initial"), LPMB_ERR_STATE, "cg_precond needs radius and xyz_initial");
    const int N = c->N, Np = c->Np;
    const double q = 2.0 * param(c, "radius");
    const double *x0 = fptr<double>(c, "xyz_initial");
    std::vector<double> hx((size_t)3 * Np);
    LPMB_D2H(c, hx.data(), x0, hx.size() * 8);
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = 1e300, hi[k] = -1e300;
        for (int i = 0; i < N; i++) {
            lo[k] = std::min(lo[k], hx[(size_t)k * Np + i]);
            hi[k] = std::max(hi[k], hx[(size_t)k * Np + i]);
        }
    }
    const int nx = (int)llround((hi[0] - lo[0]) / q) + 1, ny = (int)llround((hi[1] - lo[1]) / q) + 1, nz = (int)llround((hi[2] - lo[2]) / q) + 1;
    LPMB_REQUIRE((long long)nx * ny * nz == N, LPMB_ERR_UNSUPPORTED, "cg_precond: %d particles are not a full %d x %d x %d simple-cubic block", N, nx, ny,
                 nz);
    LPMB_REQUIRE(nx >= 5 && ny >= 5 && nz >= 5, LPMB_ERR_UNSUPPORTED, "cg_precond: the block needs at least 5 sites per edge");
    int *d_bad;
    LPMB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    LPMB_MEMSET(c, d_bad, 0, sizeof(int));
    mg_check_order_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(N, Np, x0, lo[0], lo[1], lo[2], q, nx, ny, d_bad);
    LPMB_LAUNCH_CHECK(c);
    int bad = 0;
    LPMB_D2H(c, &bad, d_bad, sizeof(int));
    cudaFree(d_bad);
    LPMB_REQUIRE(!bad, LPMB_ERR_UNSUPPORTED, "cg_precond: particles are not numbered x-fastest on an axis-aligned lattice of spacing %g", q);
    // ---- the levels: dimensions, and in slab runs this rank's part of every level (lpmb_mg_slab_plan below)
    const long long lay0 = (long long)nx * ny;
    const int W = c->world;
    M.dist = W > 1 && param(c, "mg_dist", 1.0) != 0.0;
    std::vector<long long> owned(1, nz);
    int ghost_lo = 0;
    if (M.dist) {
        LPMB_REQUIRE(W <= LPMB_PEER_MAXW, LPMB_ERR_UNSUPPORTED, "cg_precond on slabs: at most %d ranks", LPMB_PEER_MAXW);
        const int o0 = lpmb_own0(c), o1 = lpmb_own1(c);
        LPMB_REQUIRE(o0 % lay0 == 0 && o1 % lay0 == 0, LPMB_ERR_UNSUPPORTED, "cg_precond on slabs: the owned range is not whole lattice layers");
        LPMB_TRY(lpmb_ensure_staging(c, (size_t)(W + 1) * sizeof(long long)));
        long long mine = (o1 - o0) / lay0, *d_mine = (long long *)c->staging, *d_all = d_mine + 1;
        LPMB_H2D(c, d_mine, &mine, sizeof(mine));
        LPMB_TRY(lpmb_dist_allgather_bytes(c, d_mine, d_all, sizeof(mine)));
        owned.resize(W);
        LPMB_D2H(c, owned.data(), d_all, (size_t)W * sizeof(long long));
        ghost_lo = (int)(o0 / lay0);
        // the ghost exchange of level 0 writes two layers either side of the owned ones
        LPMB_REQUIRE((c->rank == 0 || ghost_lo >= 2) && (c->rank == W - 1 || nz - (int)(o1 / lay0) >= 2), LPMB_ERR_UNSUPPORTED,
                     "cg_precond on slabs: fewer than 2 ghost layers towards a neighbouring rank");
    }
    int plan[MG_MAXLEV * 8];
    LPMB_TRY(lpmb_mg_slab_plan(M.dist ? W : 1, M.dist ? c->rank : 0, owned.data(), nx, ny, nz, ghost_lo, MG_MAXLEV, plan, M.gat_off, M.gat_cnt, &M.nlev,
                               &M.lrep));
    double scale = 1.0;
    for (int l = 0; l < M.nlev; l++, scale *= 2.0) {
        MGLevel &L = M.lev[l];
        const int *P = plan + 8 * l;
        L.dist = P[0] != 0;
        L.nx = P[1], L.ny = P[2], L.nz = P[3], L.gz0 = P[4], L.oz0 = P[5], L.oz1 = P[6], L.nzg = P[7];
        L.n = (long long)L.nx * L.ny * L.nz;
        L.stride = l == 0 ? Np : L.n;
        L.scale = scale;
        LPMB_CUDA(cudaMalloc(&L.dinv, 729 * 9 * sizeof(double)));
        LPMB_CUDA(cudaMalloc(&L.u2, (size_t)3 * L.stride * 8));
        LPMB_CUDA(cudaMalloc(&L.res, (size_t)3 * L.stride * 8));
        LPMB_MEMSET(c, L.u2, 0, (size_t)3 * L.stride * 8);
        LPMB_MEMSET(c, L.res, 0, (size_t)3 * L.stride * 8);
        if (l > 0) {
            LPMB_CUDA(cudaMalloc(&L.mask, (size_t)3 * L.stride * 8));
            LPMB_CUDA(cudaMalloc(&L.u, (size_t)3 * L.stride * 8));
            LPMB_CUDA(cudaMalloc(&L.f, (size_t)3 * L.stride * 8));
            LPMB_MEMSET(c, L.mask, 0, (size_t)3 * L.stride * 8);
            LPMB_MEMSET(c, L.u, 0, (size_t)3 * L.stride * 8);
            LPMB_MEMSET(c, L.f, 0, (size_t)3 * L.stride * 8);
        }
    }
    return LPMB_OK;
}

// The stencil lives in __constant__ memory, i.e. ONCE per device: the context that uploaded last is remembered, and a
// solve of any other context (a second lattice in the same process) uploads its own copy first.

static int mg_upload_constants(lpmb_ctx *c, MGState &M)
{
    LPMB_CUDA(cudaMemcpyToSymbolAsync(c_mg_off, M.off, sizeof(M.off), 0, cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaMemcpyToSymbolAsync(c_mg_S, M.S, sizeof(M.S), 0, cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaMemcpyToSymbolAsync(c_mg_noff, &M.noff, sizeof(int), 0, cudaMemcpyHostToDevice, c->stream));
    for (int k = 0; k < MG_MAXOFF; k++)
        M.delta[k] = k < M.noff ? M.off[k][0] + MG_SX * (M.off[k][1] + MG_SY * M.off[k][2]) : 0;
    LPMB_CUDA(cudaMemcpyToSymbolAsync(c_mg_delta, M.delta, sizeof(M.delta), 0, cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));   // the host copies may change before an asynchronous upload has read them
    if (c->device >= 0 && c->device < 64)
        g_mg_const_owner[c->device] = c;
    return LPMB_OK;
}

// stencil = the off-diagonal blocks of the block row of the centre particle (re-read after every assembly)
static int mg_read_stencil(lpmb_ctx *c, MGState &M)
{
    SellMatrix &K = c->K;
    const MGLevel &L0 = M.lev[0];
    const int ic = (L0.nx / 2) + L0.nx * ((L0.ny / 2) + L0.ny * (L0.nz / 2));
    int nbc = 0;
    long long ka = 0;
    LPMB_D2H(c, &nbc, K.nbc + ic, sizeof(int));
    LPMB_D2H(c, &ka, K.sptr + (ic >> 5), sizeof(long long));
    LPMB_REQUIRE(nbc > 1 && nbc <= MG_MAXOFF, LPMB_ERR_UNSUPPORTED, "cg_precond: centre particle has %d conn entries", nbc);
    std::vector<int> col(nbc);
    std::vector<double> val((size_t)nbc * 9);
    const int lane = ic & 31;
    LPMB_CUDA(cudaMemcpy2DAsync(col.data(), sizeof(int), K.col + ka * 32 + lane, 32 * sizeof(int), sizeof(int), nbc, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaMemcpy2DAsync(val.data(), sizeof(double), K.val + ka * 9 * 32 + lane, 32 * sizeof(double), sizeof(double), (size_t)nbc * 9,
                                cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    M.noff = 0;
    const int cx = L0.nx / 2, cy = L0.ny / 2, cz = L0.nz / 2;
    for (int k = 0; k < nbc; k++) {
        const int j = col[k];
        if (j == ic)
            continue;
        const int jx = j % L0.nx, jy = (j / L0.nx) % L0.ny, jz = j / (L0.nx * L0.ny);
        M.off[M.noff][0] = jx - cx, M.off[M.noff][1] = jy - cy, M.off[M.noff][2] = jz - cz;
        for (int e = 0; e < 9; e++)
            M.S[M.noff][e] = val[(size_t)k * 9 + e];
        M.noff++;
    }
    if (M.dist) {
        // one operator for all ranks: the stencil the middle rank read (a symmetric preconditioner needs it; the ranks' own
        // interior rows differ by the local strain)
        const size_t bytes = sizeof(M.S) + sizeof(M.off) + sizeof(int);
        std::vector<char> mine(bytes), all(bytes * c->world);
        memcpy(mine.data(), M.S, sizeof(M.S));
        memcpy(mine.data() + sizeof(M.S), M.off, sizeof(M.off));
        memcpy(mine.data() + sizeof(M.S) + sizeof(M.off), &M.noff, sizeof(int));
        void *d_send = nullptr, *d_recv = nullptr;
        LPMB_CUDA(cudaMalloc(&d_send, bytes));
        LPMB_CUDA(cudaMalloc(&d_recv, bytes * c->world));
        LPMB_H2D(c, d_send, mine.data(), bytes);
        const int rc = lpmb_dist_allgather_bytes(c, d_send, d_recv, bytes);
        if (rc == LPMB_OK)
            LPMB_D2H(c, all.data(), d_recv, all.size());
        cudaFree(d_send);
        cudaFree(d_recv);
        LPMB_TRY(rc);
        const char *src = all.data() + bytes * (c->world / 2);
        memcpy(M.S, src, sizeof(M.S));
        memcpy(M.off, src + sizeof(M.S), sizeof(M.off));
        memcpy(&M.noff, src + sizeof(M.S) + sizeof(M.off), sizeof(int));
    }
    LPMB_TRY(mg_upload_constants(c, M));
    // inverse diagonal blocks per boundary class and level: D = -scale * sum of the present off-diagonal blocks
    std::vector<double> tab((size_t)729 * 9);
    for (int l = 0; l < M.nlev; l++) {
        MGLevel &L = M.lev[l];
        for (int cls = 0; cls < 729; cls++) {
            const int cxs = cls % 9, cys = (cls / 9) % 9, czs = cls / 81;
            const int lo3[3] = {cxs % 3, cys % 3, czs % 3}, hi3[3] = {cxs / 3, cys / 3, czs / 3};
            double D[9] = {0};
            for (int k = 0; k < M.noff; k++) {
                bool present = true;
                for (int a = 0; a < 3; a++)
                    present = present && M.off[k][a] >= -lo3[a] && M.off[k][a] <= hi3[a];
                if (present)
                    for (int e = 0; e < 9; e++)
                        D[e] -= L.scale * M.S[k][e];
            }
            double inv[9] = {0};
            if (!invert3(D, inv))
                for (int e = 0; e < 9; e++)
                    inv[e] = 0.0;   // a class that cannot occur on this level (isolated site)
            for (int e = 0; e < 9; e++)
                tab[(size_t)cls * 9 + e] = inv[e];
        }
        LPMB_H2D(c, L.dinv, tab.data(), tab.size() * sizeof(double));
    }
    M.ready = true;
    return LPMB_OK;
}

// hierarchy + stencil ready; coarse masks follow the solve's current DoF mask
// mask of the boundary conditions alone: 0 on constrained DoFs (boundary.c:182,214,247), 1 elsewhere
__global__ void mg_phys_mask_kernel(const int *__restrict__ bc, const int *__restrict__ fix, double *__restrict__ mask, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        mask[i] = (bc[i] == 0 || fix[i] == 0) ? 0.0 : 1.0;
}

// slab runs: refresh the two ghost layers either side of the owned layers of a distributed level's vector
static int mg_exchange(lpmb_ctx *c, const MGLevel &L, double *v)
{
    if (!L.dist)
        return LPMB_OK;
    const long long lay = (long long)L.nx * L.ny;
    return lpmb_dist_neighbor_doubles(c, v, L.stride, 3, L.oz0 * lay, (L.oz0 - 2) * lay, (L.oz1 - 2) * lay, L.oz1 * lay, 2 * lay);
}

// a vector of level l was just formed from level l - 1 for the sites this rank owns: complete it (ghost layers of a
// distributed level; the other ranks' layers of the first replicated level)
static int mg_sync_coarse(lpmb_ctx *c, MGState &M, int l, double *v)
{
    if (!M.dist)
        return LPMB_OK;
    const MGLevel &C = M.lev[l];
    if (C.dist)
        return mg_exchange(c, C, v);
    if (l == M.lrep)
        return lpmb_dist_allgatherv_doubles(c, v, C.stride, 3, M.gat_off, M.gat_cnt);
    return LPMB_OK;
}

int lpmb_mg_prepare(lpmb_ctx *c, const double *mask0)
{
    {
        auto it = g_mg.find(c);   // mg_dist switched since the hierarchy was built: build the other one
        if (it != g_mg.end() && it->second.nlev > 0 && it->second.dist != (c->world > 1 && param(c, "mg_dist", 1.0) != 0.0))
            lpmb_mg_release(c);
    }
    MGState &M = g_mg[c];
    if (M.nlev == 0) {
        const int rc = mg_build_levels(c, M);
        if (rc != LPMB_OK) {
            lpmb_mg_release(c);
            return rc;
        }
    }
    LPMB_REQUIRE(c->K.values_ready, LPMB_ERR_STATE, "stiffness matrix not available");
    if (!M.ready)
        LPMB_TRY(mg_read_stencil(c, M));
    else if (c->device < 0 || c->device >= 64 || g_mg_const_owner[c->device] != c)
        LPMB_TRY(mg_upload_constants(c, M));   // another context used the constant bank since
    M.nu = std::max(1, (int)param(c, "mg_nu", 2.0));
    M.nu_coarse = std::max(1, (int)param(c, "mg_nu_coarse", 40.0));
    // two damping factors = the degree-2 Chebyshev smoother for D^-1 A on [lambda_max / 4, lambda_max], lambda_max = 2
    // (roots 1.25 -+ 0.75 cos(pi/4)): 9 instead of 11 PCG iterations at 216^3 against omega = 0.6 twice
    // (profiles/r02r_mg_sweep_n216.log); mg_omega2 defaults to mg_omega when only that one is set
    M.omega = param(c, "mg_omega", 0.56);
    M.omega2 = param(c, "mg_omega2", c->params.count("mg_omega") ? M.omega : 1.39);
    M.lev[0].mask = const_cast<double *>(mask0);
    if (M.dist) {
        // the solve's mask also zeroes the ghost DoFs; the distributed hierarchy needs the boundary conditions alone (ghost
        // sites are ordinary sites of the global lattice whose values arrive by exchange)
        const int *bc = fptr<int>(c, "dispBC_index"), *fix = fptr<int>(c, "fix_index");
        LPMB_REQUIRE(bc && fix, LPMB_ERR_STATE, "cg_precond on slabs: BC index fields missing (lpmb_set_dof_mask / lpmb_apply_*_bc)");
        const size_t n3 = (size_t)3 * c->Np;
        if (!M.mask_phys)
            LPMB_CUDA(cudaMalloc(&M.mask_phys, n3 * 8));
        mg_phys_mask_kernel<<<lpmb_blocks((long long)n3, 256), 256, 0, c->stream>>>(bc, fix, M.mask_phys, n3);
        LPMB_LAUNCH_CHECK(c);
        M.lev[0].mask = M.mask_phys;
    }
    for (int l = 1; l < M.nlev; l++) {
        const MGLevel &F = M.lev[l - 1];
        MGLevel &C = M.lev[l];
        mg_coarse_mask_kernel<<<lpmb_blocks(C.n, 128), 128, 0, c->stream>>>(F.nx, F.ny, F.nz, F.stride, C.nx, C.ny, C.nz, C.stride, F.gz0, C.gz0, C.nzg,
                                                                            F.mask, C.mask);
        LPMB_LAUNCH_CHECK(c);
        LPMB_TRY(mg_sync_coarse(c, M, l, C.mask));   // slab runs: ghost layers / the other ranks' parts
    }
    return LPMB_OK;
}

// nu damped-Jacobi sweeps on level l starting from u = 0 (first = true) or from L.u; result in L.u
static int mg_smooth(lpmb_ctx *c, MGState &M, int l, int nu, bool first, const double *done, bool ghosts_valid = false)
{
    MGLevel &L = M.lev[l];
    for (int s = 0; s < nu; s++) {
        const double om = (s & 1) ? M.omega2 : M.omega;
        if (first && s == 0) {
            LPMB_TRY(mg_launch_stencil<2>(c, L, nullptr, L.f, L.u2, om, done));
        } else {
            if (!(ghosts_valid && s == 0))
                LPMB_TRY(mg_exchange(c, L, L.u));   // slab runs: the stencil reaches two layers into the neighbours' sites
            LPMB_TRY(mg_launch_stencil<0>(c, L, L.u, L.f, L.u2, om, done));
        }
        std::swap(L.u, L.u2);
    }
    return LPMB_OK;
}

static int mg_vcycle(lpmb_ctx *c, MGState &M, int l, const double *done)
{
    MGLevel &L = M.lev[l];
    if (l == M.nlev - 1) {
        if (L.n <= MG_COARSE_MAX && param(c, "mg_coarse_fused", 1.0) != 0.0) {
            mg_coarse_solve_kernel<<<1, 256, 0, c->stream>>>(L.nx, L.ny, L.nz, L.stride, L.scale, L.dinv, L.mask, L.f, L.u, M.omega, M.nu_coarse, done);
            LPMB_LAUNCH_CHECK(c);
            return LPMB_OK;
        }
        return mg_smooth(c, M, l, M.nu_coarse, true, done);
    }
    LPMB_TRY(mg_smooth(c, M, l, M.nu, true, done));
    LPMB_TRY(mg_exchange(c, L, L.u));
    LPMB_TRY(mg_launch_stencil<1>(c, L, L.u, L.f, L.res, 0.0, done));
    MGLevel &C = M.lev[l + 1];
    LPMB_TRY(mg_exchange(c, L, L.res));   // full weighting reaches one layer into the neighbours' sites
    mg_restrict_kernel<<<lpmb_blocks(C.n, 128), 128, 0, c->stream>>>(L.nx, L.ny, L.nz, L.stride, C.nx, C.ny, C.nz, C.stride, L.gz0, C.gz0, C.nzg, L.res,
                                                                     C.mask, C.f, done);
    LPMB_LAUNCH_CHECK(c);
    if (M.dist && !C.dist && l + 1 == M.lrep)
        LPMB_TRY(mg_sync_coarse(c, M, l + 1, C.f));   // first replicated level: every rank gets the whole right-hand side
    LPMB_TRY(mg_vcycle(c, M, l + 1, done));
    LPMB_TRY(mg_exchange(c, C, C.u));   // the interpolation reads one coarse layer beyond the owned ones
    mg_prolong_kernel<<<lpmb_blocks(L.n, 128), 128, 0, c->stream>>>(L.nx, L.ny, L.nz, L.stride, C.nx, C.ny, C.nz, C.stride, L.gz0, C.gz0, C.nzg, C.u,
                                                                    L.mask, L.u, done);
    LPMB_LAUNCH_CHECK(c);
    // The interpolation ran on every local site: the two ghost layers of u (valid since the exchange before the residual)
    // received the same correction as on their owner (it read the coarse ghost layers just exchanged, and the hierarchy's
    // level-0 mask is the boundary-condition mask, the same on both sides) -- the first post-smoothing sweep needs no exchange.
    return mg_smooth(c, M, l, M.nu, false, done, M.dist && L.dist);
}

// z = V-cycle(r) on the context's original-order vectors ([3][Np]); r is not modified.  `done` (device, may be null):
// every kernel returns at once when *done != 0 (the PCG loop issues iterations in batches).
int lpmb_mg_apply(lpmb_ctx *c, const double *r, double *z, const double *done)
{
    MGState &M = g_mg[c];
    LPMB_REQUIRE(M.ready && M.nlev > 0, LPMB_ERR_STATE, "multigrid hierarchy not prepared");
    MGLevel &L0 = M.lev[0];
    // level 0 smooths between z and its scratch twin; an even total number of swaps leaves the result in z
    L0.f = const_cast<double *>(r);
    L0.u = z;
    LPMB_TRY(mg_vcycle(c, M, 0, done));
    if (L0.u != z) {   // odd number of sweeps in total: the result sits in the scratch buffer
        LPMB_CUDA(cudaMemcpyAsync(z, L0.u, (size_t)3 * L0.stride * 8, cudaMemcpyDeviceToDevice, c->stream));
        L0.u2 = L0.u;
    }
    L0.u = nullptr;
    return LPMB_OK;
}

int lpmb_mg_levels(lpmb_ctx *c)
{
    auto it = g_mg.find(c);
    return it == g_mg.end() ? 0 : it->second.nlev;
}
