// lpmb_brick.cu -- brick-blocked symmetric SpMV: every stored block is fetched from HBM once and used for
// BOTH of its contributions on the same SM.
//
// Why.  K is symmetric (stiffness.c:441-481 keeps one triangle); the default SELL kernel streams both triangles
// (61 blocks * 76 B per interior SC particle) and already runs at the HBM roofline, so the only way to make a CG
// iteration faster is to move fewer bytes.  Re-using a block through L2 does not work on this chip (measured in
// round 1: 0.85x of the full-format kernel at 216^3); re-using it inside one CTA does.
//
// How.  Particles are regrouped (internally only; the ABI keeps the reference's numbering) into bricks of 8x8x8
// lattice sites = 512 rows = one CTA of 512 threads, thread r <-> row r = local lattice position.  A pair {i,j}
// is owned by the end point from which the displacement j-i lies in the positive half space (dz>0, or dz=0 &
// dy>0, or dz=dy=0 & dx>0), so ownership does not depend on any numbering.  Displacements are grouped in classes
// (30 positive ones + the diagonal for the simple-cubic 2-hop stencil); the brick stores, class-major, then z-layer-major and
// coalesced over the 64 sites of a layer, only the 3x3 blocks K_{i,i+d} -- no column indices: the neighbour of row r in class u sits
// at a fixed offset in the brick's 12 x 12 x 10 extended box, absent blocks are zeros.  For each class, thread r
// takes its block from the shared-memory tile TMA delivered and forms
//        own[r]      += K   x_j        (registers; x of the extended box is staged in shared memory)
//        acc[i+d]    += K^T x_i        (shared memory, same slot index as x_j)
// A class is a translation, so within a class all targets are distinct: plain shared-memory read-modify-write,
// one __syncthreads per class, fixed summation order -> deterministic, no atomics.  Contributions that land
// outside the brick are written to a per-brick staging box and folded in by a second, small gather kernel that
// adds the (at most 7) source bricks of a row in a fixed order; it also applies the DoF mask and produces the
// p.Ap partials.  Traffic per interior SC particle: 31*72 B of matrix + ~90 B staging + vectors = 2.4 KB instead
// of 4.69 KB; measured 3.84 ms per SpMV at 216^3 against 6.97 ms for the full format.
//
// Scope: axis-aligned simple-cubic lattices in 3-D (checked on the device; anything else keeps the full-format
// kernel), one GPU or z-slabs (rows = owned layers +- the 2-layer CG halo; halo rows of the search direction are
// pushed by the neighbours through peer memory, lpmb_peer.cu, or exchanged with NCCL).  Enabled explicitly with
// lpmb_matrix_enable_bricks(); the full-format values stay the master copy and are mirrored after every assembly.
#include <algorithm>

#include "lpmb_internal.cuh"

#define BE 8               // brick edge (lattice sites)
#define BR 512             // rows per brick = BE^3 = threads per CTA
#define EXX 12             // extended box: x,y in [-2, BE+2), z in [0, BE+2)
#define EXY 12
#define EXZ 10
#define NSLOT (EXX * EXY * EXZ)
#define MAXCLS 32

struct BrickMatrix {
    bool enabled = false, pattern_ready = false, values_ready = false;
    int nbx = 0, nby = 0, nbz = 0, nbricks = 0, ncls = 0;
    long long P = 0;                 // permuted, padded vector length per component = nbricks * BR
    double q = 0, ox = 0, oy = 0, oz = 0;
    int *perm = nullptr;             // [P] permuted row -> original particle (-1 = padding)
    int *inv = nullptr;              // [Np] original particle -> permuted row
    int *ic = nullptr;               // [3][Np] integer lattice coordinates
    double *bval = nullptr;          // [nbricks][ncls][8 z-layers][9][64]
    double *stage = nullptr;         // [nbricks][3][NSLOT]
    double *ypart = nullptr;         // [3][P]
    double *r = nullptr, *p = nullptr, *ap = nullptr, *x = nullptr, *b = nullptr, *mask = nullptr;  // CG vectors [3][P]
    int cls_d[MAXCLS][3];            // displacement of each class (class 0 = diagonal)
    int key2cls[125];
    // rows a class tile really needs, per (brick layer, class): local z-layers [zr[..][0], zr[..][1]) of the 8.  A row is
    // needed when it exists and it, or its partner i+d, is OWNED: the top rows of a partial brick layer are empty, the
    // upper CG-halo rows of a slab own no pair that touches an owned row (pairs are owned by their lower end), the lower
    // halo rows only through classes that reach up into the slab.  Only that part of a tile is streamed from HBM.
    unsigned char *zr = nullptr;     // [nbz][ncls][2]
    std::vector<unsigned char> h_zr;
    int nz = 0, own_z0 = 0, own_z1 = 0;
};

static std::map<lpmb_ctx *, BrickMatrix> g_bricks;

__constant__ int c_cls_off[MAXCLS];  // slot offset dx + EXX (dy + EXY dz) of each class
__constant__ int c_key2cls[125];

void lpmb_brick_release(lpmb_ctx *c)
{
    auto it = g_bricks.find(c);
    if (it == g_bricks.end())
        return;
    BrickMatrix &B = it->second;
    lpmb_peer_halo_release(c);
    cudaFree(B.perm); cudaFree(B.inv); cudaFree(B.ic); cudaFree(B.bval); cudaFree(B.stage); cudaFree(B.ypart); cudaFree(B.zr);
    cudaFree(B.r); cudaFree(B.p); cudaFree(B.ap); cudaFree(B.x); cudaFree(B.b); cudaFree(B.mask);
    g_bricks.erase(it);
}

void lpmb_brick_touch(lpmb_ctx *c)   // K.val changed: every mirror of it is stale
{
    c->K.rows_ready = false;
    lpmb_mg_touch(c);
    auto it = g_bricks.find(c);
    if (it != g_bricks.end())
        it->second.values_ready = false;
}

bool lpmb_brick_active(lpmb_ctx *c)
{
    auto it = g_bricks.find(c);
    return it != g_bricks.end() && it->second.enabled;
}

// ---- set-up kernels ------------------------------------------------------------------------------
// integer lattice coordinates of every particle; max deviation from the lattice (alignment check)
__global__ void brick_quantize_kernel(int N, int Np, const double *__restrict__ x0, double ox, double oy, double oz, double q, int nzdom,
                                      int *__restrict__ ic /* [3][Np] */, double *__restrict__ maxdev)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double dev = 0.0;
    if (i < N) {
        const double fx = (x0[i] - ox) / q, fy = (x0[(size_t)Np + i] - oy) / q, fz = (x0[(size_t)2 * Np + i] - oz) / q;
        const double rx = rint(fx), ry = rint(fy), rz = rint(fz);
        const bool inside = rz >= 0.0 && rz < (double)nzdom;  // slab runs: ghost layers beyond the CG halo carry no row
        ic[i] = (int)rx;
        ic[(size_t)Np + i] = (int)ry;
        ic[(size_t)2 * Np + i] = inside ? (int)rz : -1;
        dev = fmax(fabs(fx - rx), fmax(fabs(fy - ry), fabs(fz - rz)));
    }
    // warp max, then one atomicMax on the bit pattern (non-negative doubles order like unsigned integers)
    for (int o = 16; o > 0; o >>= 1)
        dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long *>(maxdev), (unsigned long long)__double_as_longlong(dev));
}

// row of particle i inside its brick = its local lattice position (bricks are addressed geometrically, so the
// permutation is a pure function of the coordinates: deterministic)
__global__ void brick_place_kernel(int N, int Np, const int *__restrict__ ic, int nbx, int nby, int *__restrict__ perm, int *__restrict__ inv,
                                   int *__restrict__ clash)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const int ix = ic[i], iy = ic[(size_t)Np + i], iz = ic[(size_t)2 * Np + i];
    if (iz < 0) {
        inv[i] = -1;
        return;
    }
    const int b = (ix / BE) + nbx * ((iy / BE) + nby * (iz / BE));
    const int lx = ix % BE, ly = iy % BE, lz = iz % BE;
    const int r = lx + BE * (ly + BE * lz);
    const long long prow = (long long)b * BR + r;
    if (atomicCAS(&perm[prow], -1, i) != -1)
        atomicExch(clash, 1);  // two particles on one lattice site
    inv[i] = (int)prow;
}

// which displacement keys occur (key = (dx+2) + 5 (dy+2) + 25 (dz+2)); flags[125], flags[125] = out-of-reach seen
__global__ void brick_keys_kernel(int N, int Np, const int *__restrict__ ic, const long long *__restrict__ sptr, const int *__restrict__ col,
                                  const int *__restrict__ nbc, int *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const long long base = sptr[i >> 5] * 32 + (i & 31);
    if (ic[(size_t)2 * Np + i] < 0)
        return;
    for (int k = 0; k < nbc[i]; k++) {
        const int j = col[base + (long long)k * 32];
        if (ic[(size_t)2 * Np + j] < 0)
            continue;
        const int dx = ic[j] - ic[i], dy = ic[(size_t)Np + j] - ic[(size_t)Np + i], dz = ic[(size_t)2 * Np + j] - ic[(size_t)2 * Np + i];
        if (dx < -2 || dx > 2 || dy < -2 || dy > 2 || dz < -2 || dz > 2)
            flags[125] = 1;
        else
            flags[(dx + 2) + 5 * (dy + 2) + 25 * (dz + 2)] = 1;
    }
}

__global__ void brick_fill_kernel(int N, int Np, const int *__restrict__ ic, const int *__restrict__ inv, const long long *__restrict__ sptr,
                                  const int *__restrict__ col, const double *__restrict__ val, const int *__restrict__ nbc, int ncls,
                                  double *__restrict__ bval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const long long prow = inv[i];
    if (prow < 0)
        return;
    const long long b = prow / BR;
    const int r = (int)(prow % BR);
    const long long kbase = sptr[i >> 5];
    const int lane = i & 31;
    for (int k = 0; k < nbc[i]; k++) {
        const int j = col[(kbase + k) * 32 + lane];
        if (inv[j] < 0)
            continue;  // neighbour outside the slab's CG halo
        const int dx = ic[j] - ic[i], dy = ic[(size_t)Np + j] - ic[(size_t)Np + i], dz = ic[(size_t)2 * Np + j] - ic[(size_t)2 * Np + i];
        const int u = c_key2cls[(dx + 2) + 5 * (dy + 2) + 25 * (dz + 2)];
        if (u < 0)
            continue;  // negative half space: owned by the other end point
#pragma unroll
        for (int e = 0; e < 9; e++)
            bval[(((b * ncls + u) * BE + (r >> 6)) * 9 + e) * 64 + (r & 63)] = val[((kbase + k) * 9 + e) * 32 + lane];
    }
}

// original-order [3][Np] <-> permuted [3][P]
__global__ void brick_to_perm_kernel(long long P, int Np, const int *__restrict__ perm, const double *__restrict__ src, double *__restrict__ dst,
                                     double fill)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * P)
        return;
    const int comp = (int)(t / P);
    const long long prow = t % P;
    const int i = perm[prow];
    dst[t] = i >= 0 ? src[(size_t)comp * Np + i] : fill;
}

__global__ void brick_from_perm_kernel(int N, int Np, long long P, const int *__restrict__ inv, const double *__restrict__ src, double *__restrict__ dst)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3LL * N)
        return;
    const int comp = (int)(t / N), i = (int)(t % N);
    dst[(size_t)comp * Np + i] = inv[i] >= 0 ? src[(size_t)comp * P + inv[i]] : 0.0;
}

// halo rows only: [i0, i0+count) of the original numbering, permuted <-> original (slab runs)
__global__ void brick_range_kernel(int i0, int count, int Np, long long P, const int *__restrict__ inv, double *__restrict__ perm_vec,
                                   double *__restrict__ orig_vec, int to_perm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * count)
        return;
    const int comp = t / count, i = i0 + t % count;
    const long long prow = inv[i];
    if (to_perm) {
        if (prow >= 0)
            perm_vec[(size_t)comp * P + prow] = orig_vec[(size_t)comp * Np + i];
    } else {
        orig_vec[(size_t)comp * Np + i] = prow >= 0 ? perm_vec[(size_t)comp * P + prow] : 0.0;
    }
}

// ---- the SpMV ------------------------------------------------------------------------------------
// One persistent CTA per SM walks over bricks.  The 36 KB class tiles ([8 z-layers][9][64] doubles, contiguous in HBM;
// only the z-layers whose rows are needed -- `zr` table -- are copied: one contiguous run) are
// streamed into a 4-deep shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier transaction counts),
// issued by one thread and running ahead across brick boundaries, so ~147 KB per SM are in flight regardless of
// the per-class barrier.  x of the brick's extended box lives in shared memory too: x_j and the transposed
// target share the same slot index, so the kernel needs no column indices at all (absent blocks are zero).
#define NST 4
#define TILE_DOUBLES (9 * BR)
#define TILE_BYTES (TILE_DOUBLES * 8)

struct BrickSmem {
    double tile[NST][BE][9][64];   // z-layer major: the needed z-range of a tile is one contiguous run
    double acc[3][NSLOT];
    double xs[3][NSLOT];
    unsigned long long full[NST];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// LAZY (slab runs with the peer-memory halo push, param brick_lazy_wait; experimental, off by default): a CTA does not wait
// for the neighbours' pushes up front but only before the first brick whose extended box really holds halo rows of x
// -- the lower ghost layers sit in brick layer 0, the upper ones are reached from the top brick layer(s) -- and the
// bricks are walked in an order rotated by `rot` (one brick layer), so that layer 0 comes last: the pushes land while the
// interior layers are being processed.  LAZY = false is the kernel as measured in round 1 (wait up front, natural order).
template <bool LAZY>
__global__ void __launch_bounds__(BR, 1)
brick_spmv_kernel(int nbricks, int ncls, int nbx, int nby, int nbz, long long P, const double *__restrict__ bval,
                  const unsigned char *__restrict__ zr, const double *__restrict__ x, double *__restrict__ ypart, double *__restrict__ stage,
                  const double *__restrict__ scal, PeerWait halo_wait, int rot, int own_z0, int own_z1)
{
    extern __shared__ __align__(128) unsigned char brick_smem_raw[];
    BrickSmem &S = *reinterpret_cast<BrickSmem *>(brick_smem_raw);
    if (scal && scal[7] != 0.0)  // S_DONE
        return;
    if (!LAZY)
        lpmb_peer_wait(halo_wait);  // slab runs: the neighbours' pushes of the halo rows of x have landed (lpmb_peer.cu)
    bool waited_lo = !LAZY || halo_wait.n == 0, waited_hi = waited_lo;
    // logical position in this CTA's walk -> brick
    auto phys = [&](long long k) -> long long {
        if (!LAZY)
            return k;
        const long long b = k + rot;
        return b >= nbricks ? b - nbricks : b;
    };
    const int r = threadIdx.x;
    const int lx = r & 7, ly = (r >> 3) & 7, lz = r >> 6;
    const int myslot = (lx + 2) + EXX * ((ly + 2) + EXY * lz);
    const int nloc = (int)blockIdx.x < nbricks ? (nbricks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const long long ntiles = (long long)nloc * ncls;
    if (r == 0) {
        for (int s = 0; s < NST; s++)
            mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long t) {
        const int s = (int)(t % NST);
        const long long brick = phys(blockIdx.x + (t / ncls) * (long long)gridDim.x);
        const int u = (int)(t % ncls);
        const int bzl = (int)(brick / ((long long)nbx * nby));
        const int z0 = zr[(bzl * ncls + u) * 2], z1 = zr[(bzl * ncls + u) * 2 + 1];
        const double *src = bval + (brick * ncls + u) * (long long)TILE_DOUBLES;
        // only the z-layers [z0, z1) (nothing at all if empty): one bulk copy of 4 608 bytes per layer
        const unsigned n = z1 > z0 ? (unsigned)((z1 - z0) * 9 * 64 * 8) : 0u;
        mbar_expect_tx(&S.full[s], n);
        if (n)
            bulk_g2s(&S.tile[s][z0][0][0], src + z0 * 9 * 64, n, &S.full[s]);
    };
    if (r == 0)
        for (long long t = 0; t < NST && t < ntiles; t++)
            issue(t);
    long long t = 0;
    for (int lb = 0; lb < nloc; lb++) {
        const long long brick = phys(blockIdx.x + (long long)lb * gridDim.x);
        const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((long long)nbx * nby));
        if (LAZY) {
            // uniform over the CTA (depends on bz only): flag 0 is raised by rank-1 (lower ghosts), flag 1 by rank+1
            const bool need_lo = !waited_lo && bz * BE < own_z0, need_hi = !waited_hi && bz * BE + EXZ > own_z1;
            if (need_lo || need_hi) {
                if (r == 0) {
                    if (need_lo)
                        while (lpmb_ld_acquire_sys(halo_wait.seqs + 0) < halo_wait.seq) {
                        }
                    if (need_hi)
                        while (lpmb_ld_acquire_sys(halo_wait.seqs + 1) < halo_wait.seq) {
                        }
                }
                __syncthreads();
                waited_lo |= need_lo;
                waited_hi |= need_hi;
            }
        }
        // x over the extended box (zero outside the lattice), accumulators to zero
        for (int s = r; s < NSLOT; s += BR) {
            const int gx = bx * BE + (s % EXX) - 2, gy = by * BE + ((s / EXX) % EXY) - 2, gz = bz * BE + s / (EXX * EXY);
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (gx >= 0 && gx < nbx * BE && gy >= 0 && gy < nby * BE && gz < nbz * BE) {
                const long long src = ((gx >> 3) + (long long)nbx * ((gy >> 3) + (long long)nby * (gz >> 3))) * BR + ((gx & 7) + BE * ((gy & 7) + BE * (gz & 7)));
                v0 = x[src];
                v1 = x[P + src];
                v2 = x[2 * P + src];
            }
            S.xs[0][s] = v0;
            S.xs[1][s] = v1;
            S.xs[2][s] = v2;
            S.acc[0][s] = S.acc[1][s] = S.acc[2][s] = 0.0;
        }
        __syncthreads();
        const double xi0 = S.xs[0][myslot], xi1 = S.xs[1][myslot], xi2 = S.xs[2][myslot];
        double o0 = 0.0, o1 = 0.0, o2 = 0.0;
        const unsigned char *zrb = zr + (size_t)bz * ncls * 2;
        for (int u = 0; u < ncls; u++, t++) {
            const int st = (int)(t % NST);
            mbar_wait(&S.full[st], (unsigned)((t / NST) & 1));
            if (lz >= zrb[2 * u] && lz < zrb[2 * u + 1]) {  // rows outside the streamed z-range: not needed, tile part stale
                double a[9];
#pragma unroll
                for (int e = 0; e < 9; e++)
                    a[e] = S.tile[st][lz][e][r & 63];
                const int slot = myslot + c_cls_off[u];
                const double xj0 = S.xs[0][slot], xj1 = S.xs[1][slot], xj2 = S.xs[2][slot];
                o0 = fma(a[0], xj0, fma(a[1], xj1, fma(a[2], xj2, o0)));
                o1 = fma(a[3], xj0, fma(a[4], xj1, fma(a[5], xj2, o1)));
                o2 = fma(a[6], xj0, fma(a[7], xj1, fma(a[8], xj2, o2)));
                if (u > 0) {  // class 0 is the diagonal block: no transposed partner
                    S.acc[0][slot] += fma(a[0], xi0, fma(a[3], xi1, a[6] * xi2));
                    S.acc[1][slot] += fma(a[1], xi0, fma(a[4], xi1, a[7] * xi2));
                    S.acc[2][slot] += fma(a[2], xi0, fma(a[5], xi1, a[8] * xi2));
                }
            }
            __syncthreads();  // tile consumed by everybody; the next class may hit the same slots
            if (r == 0 && t + NST < ntiles)
                issue(t + NST);
        }
        const long long prow = brick * BR + r;
        ypart[prow] = o0 + S.acc[0][myslot];
        ypart[P + prow] = o1 + S.acc[1][myslot];
        ypart[2 * P + prow] = o2 + S.acc[2][myslot];
        // contributions that left the brick (all written every launch: the gather kernel reads all of them)
        double *sg = stage + brick * 3 * NSLOT;
        for (int s = r; s < NSLOT; s += BR) {
            const int ex = s % EXX, ey = (s / EXX) % EXY, ez = s / (EXX * EXY);
            const bool own = ex >= 2 && ex < BE + 2 && ey >= 2 && ey < BE + 2 && ez < BE;
            if (!own) {
                sg[s] = S.acc[0][s];
                sg[NSLOT + s] = S.acc[1][s];
                sg[2 * NSLOT + s] = S.acc[2][s];
            }
        }
        __syncthreads();  // before the next brick's prologue overwrites xs / acc
    }
}

// y = mask .* (ypart + contributions staged by neighbouring bricks), partial dot products x.y
template <bool DOT>
__global__ void __launch_bounds__(256)
brick_gather_kernel(long long P, int nbx, int nby, int nbz, const double *__restrict__ ypart, const double *__restrict__ stage,
                    const double *__restrict__ mask, const double *__restrict__ x, double *__restrict__ y, double *__restrict__ partials,
                    const double *__restrict__ scal, PeerPublish pub)
{
    __shared__ double red[8];
    if (DOT && scal && scal[7] != 0.0)
        return;
    double dot = 0.0;
    for (long long prow = (long long)blockIdx.x * 256 + threadIdx.x; prow < P; prow += (long long)gridDim.x * 256) {
        const long long b = prow / BR;
        const int r = (int)(prow % BR);
        const int bx = (int)(b % nbx), by = (int)((b / nbx) % nby), bz = (int)(b / ((long long)nbx * nby));
        const int lx = r & 7, ly = (r >> 3) & 7, lz = r >> 6;
        double v0 = ypart[prow], v1 = ypart[P + prow], v2 = ypart[2 * P + prow];
        // Source bricks: a row within 2 sites of a brick face also sits in the extended box of the brick across that
        // face (z: only the brick below, because pairs are owned by their lower end point).  With one candidate
        // offset per axis (fx, fy in {-1,0,+1}, fz in {-1,0}) the sources are the 7 non-empty combinations; all
        // loads are issued unconditionally from clamped addresses (independent, no divergent latency chains) and
        // added in a fixed order.
        const int fx = lx < 2 ? -1 : (lx >= BE - 2 ? 1 : 0), fy = ly < 2 ? -1 : (ly >= BE - 2 ? 1 : 0), fz = lz < 2 ? -1 : 0;
        const bool okx = fx != 0 && bx + fx >= 0 && bx + fx < nbx, oky = fy != 0 && by + fy >= 0 && by + fy < nby, okz = fz != 0 && bz + fz >= 0;
        double c0[7], c1[7], c2[7];
#pragma unroll
        for (int k = 1; k < 8; k++) {
            const bool ux = k & 1, uy = k & 2, uz = k & 4;
            const bool ok = (!ux || okx) && (!uy || oky) && (!uz || okz);
            const int dbx = ux ? fx : 0, dby = uy ? fy : 0, dbz = uz ? fz : 0;
            const long long sb = ok ? (bx + dbx) + (long long)nbx * ((by + dby) + (long long)nby * (bz + dbz)) : b;
            const int sl = ok ? (lx - BE * dbx + 2) + EXX * ((ly - BE * dby + 2) + EXY * (lz - BE * dbz)) : 0;
            const double *sg = stage + sb * 3 * NSLOT + sl;
            const double w = ok ? 1.0 : 0.0;
            c0[k - 1] = w * sg[0];
            c1[k - 1] = w * sg[NSLOT];
            c2[k - 1] = w * sg[2 * NSLOT];
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            v0 += c0[k];
            v1 += c1[k];
            v2 += c2[k];
        }
        if (mask) {
            v0 *= mask[prow];
            v1 *= mask[P + prow];
            v2 *= mask[2 * P + prow];
        }
        if (DOT)
            dot = fma(v0, x[prow], fma(v1, x[P + prow], fma(v2, x[2 * P + prow], dot)));
        y[prow] = v0;
        y[P + prow] = v1;
        y[2 * P + prow] = v2;
    }
    if (DOT) {
        for (int o = 16; o > 0; o >>= 1)
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if ((threadIdx.x & 31) == 0)
            red[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = 0.0;
            for (int k = 0; k < 8; k++)
                tsum += red[k];
            partials[blockIdx.x] = tsum;
        }
        // slab runs on the peer-memory path: the last block folds the partials and publishes this rank's p.Ap (lpmb_peer.cu)
        if (pub.world > 0 && lpmb_last_block(pub.counter))
            lpmb_peer_publish_block(partials, gridDim.x, pub, red);
    }
}

// ---- host side -------------------------------------------------------------------------------------
static int brick_upload_tables(const BrickMatrix &B)
{
    int off[MAXCLS] = {0};
    for (int u = 0; u < B.ncls; u++)
        off[u] = B.cls_d[u][0] + EXX * (B.cls_d[u][1] + EXY * B.cls_d[u][2]);
    LPMB_CUDA(cudaMemcpyToSymbol(c_cls_off, off, sizeof(off)));
    LPMB_CUDA(cudaMemcpyToSymbol(c_key2cls, B.key2cls, sizeof(int) * 125));
    return LPMB_OK;
}

static int brick_build_pattern(lpmb_ctx *c, BrickMatrix &B)
{
    LPMB_REQUIRE(c->dim == 3 && c->lattice == LPMB_LATTICE_SC, LPMB_ERR_UNSUPPORTED, "brick SpMV: simple-cubic 3-D lattices only");
    LPMB_REQUIRE(c->params.count("radius") && c->fields.count("xyz_initial") && c->K.pattern_ready, LPMB_ERR_STATE,
                 "brick SpMV needs radius, xyz_initial and the connectivity");
    const int N = c->N, Np = c->Np;
    const double *x0 = fptr<double>(c, "xyz_initial");
    B.q = 2.0 * param(c, "radius");
    // lattice origin = component-wise minimum.  Slab runs (world > 1): rows exist only for the owned z-layers plus the
    // CG halo (2 layers); the outer, position-only ghost layers are left out.
    std::vector<double> hx((size_t)3 * Np);
    LPMB_D2H(c, hx.data(), x0, hx.size() * 8);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < N; i++) {
            lo[k] = std::min(lo[k], hx[(size_t)k * Np + i]);
            hi[k] = std::max(hi[k], hx[(size_t)k * Np + i]);
        }
    if (c->world > 1) {
        const int own0 = lpmb_own0(c), own1 = lpmb_own1(c);
        LPMB_REQUIRE(own1 > own0, LPMB_ERR_STATE, "brick SpMV: empty slab");
        double zl = 1e300, zh = -1e300;
        for (int i = own0; i < own1; i++) {
            zl = std::min(zl, hx[(size_t)2 * Np + i]);
            zh = std::max(zh, hx[(size_t)2 * Np + i]);
        }
        lo[2] = std::max(lo[2], zl - 2.0 * B.q * (1.0 + 1e-9));
        hi[2] = std::min(hi[2], zh + 2.0 * B.q * (1.0 + 1e-9));
        // snap to the lattice planes actually present
        lo[2] = zl - B.q * std::floor((zl - lo[2]) / B.q + 1e-6);
        hi[2] = zh + B.q * std::floor((hi[2] - zh) / B.q + 1e-6);
    }
    double own_zl = lo[2], own_zh = hi[2];  // z of the first / last OWNED lattice layer (single GPU: all of them)
    if (c->world > 1) {
        own_zl = 1e300, own_zh = -1e300;
        for (int i = lpmb_own0(c); i < lpmb_own1(c); i++) {
            own_zl = std::min(own_zl, hx[(size_t)2 * Np + i]);
            own_zh = std::max(own_zh, hx[(size_t)2 * Np + i]);
        }
    }
    hx.clear();
    hx.shrink_to_fit();
    B.ox = lo[0];
    B.oy = lo[1];
    B.oz = lo[2];
    const int nx = (int)llround((hi[0] - lo[0]) / B.q) + 1, ny = (int)llround((hi[1] - lo[1]) / B.q) + 1, nz = (int)llround((hi[2] - lo[2]) / B.q) + 1;
    B.nbx = (nx + BE - 1) / BE;
    B.nby = (ny + BE - 1) / BE;
    B.nbz = (nz + BE - 1) / BE;
    B.nz = nz;
    B.own_z0 = (int)llround((own_zl - lo[2]) / B.q);
    B.own_z1 = (int)llround((own_zh - lo[2]) / B.q) + 1;
    // test hook: emulate a slab on one GPU (rows outside [brick_own_z0, brick_own_z1) are then "halo": their products are
    // not formed, exactly as in a slab run -- tests/test_solver_gpu.py compares the owned rows with the full format)
    if (c->world == 1 && c->params.count("brick_own_z0") && c->params.count("brick_own_z1")) {
        B.own_z0 = std::max(0, (int)param(c, "brick_own_z0"));
        B.own_z1 = std::min(nz, (int)param(c, "brick_own_z1"));
    }
    B.nbricks = B.nbx * B.nby * B.nbz;
    B.P = (long long)B.nbricks * BR;
    LPMB_REQUIRE(B.P < (1LL << 31), LPMB_ERR_UNSUPPORTED, "brick SpMV: %lld padded rows exceed 32-bit indices", B.P);
    int *ic;
    double *d_dev;
    LPMB_CUDA(cudaMalloc(&ic, (size_t)3 * Np * sizeof(int)));
    LPMB_CUDA(cudaMalloc(&d_dev, sizeof(double)));
    LPMB_MEMSET(c, d_dev, 0, sizeof(double));
    brick_quantize_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(N, Np, x0, B.ox, B.oy, B.oz, B.q, nz, ic, d_dev);
    LPMB_LAUNCH_CHECK(c);
    double dev = 0.0;
    LPMB_CUDA(cudaMemcpyAsync(&dev, d_dev, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_dev);
    if (dev > 1e-6) {
        cudaFree(ic);
        lpmb_set_error("brick SpMV: particles are not on an axis-aligned lattice of spacing %g (max deviation %g)", B.q, dev);
        return LPMB_ERR_UNSUPPORTED;
    }
    LPMB_CUDA(cudaMalloc(&B.perm, (size_t)B.P * sizeof(int)));
    LPMB_MEMSET(c, B.perm, 0xff, (size_t)B.P * sizeof(int));
    LPMB_CUDA(cudaMalloc(&B.inv, (size_t)Np * sizeof(int)));
    int *d_flags;
    LPMB_CUDA(cudaMalloc(&d_flags, 128 * sizeof(int)));
    LPMB_MEMSET(c, d_flags, 0, 128 * sizeof(int));
    brick_place_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(N, Np, ic, B.nbx, B.nby, B.perm, B.inv, d_flags + 126);
    LPMB_LAUNCH_CHECK(c);
    brick_keys_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(N, Np, ic, c->K.sptr, c->K.col, c->K.nbc, d_flags);
    LPMB_LAUNCH_CHECK(c);
    if (c->world > 1) {
        // every owned particle and every CG-halo particle must have a row (holds for z-slabs made of whole layers)
        const int i0 = lpmb_own0(c) - (c->rank > 0 ? c->narrow_recv_lo : 0), i1 = lpmb_own1(c) + (c->rank < c->world - 1 ? c->narrow_recv_hi : 0);
        std::vector<int> hinv((size_t)(i1 - i0));
        LPMB_CUDA(cudaMemcpyAsync(hinv.data(), B.inv + i0, hinv.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        for (int v : hinv)
            if (v < 0) {
                cudaFree(ic);
                cudaFree(d_flags);
                lpmb_set_error("brick SpMV: the slab is not a stack of whole z-layers (a halo particle lies outside owned +- 2 layers)");
                return LPMB_ERR_UNSUPPORTED;
            }
    }
    int flags[128];
    LPMB_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_flags);
    if (flags[125] || flags[126]) {
        cudaFree(ic);
        lpmb_set_error(flags[126] ? "brick SpMV: two particles share a lattice site" : "brick SpMV: a conn entry reaches further than 2 lattice steps");
        return LPMB_ERR_UNSUPPORTED;
    }
    // classes: 0 = diagonal, then the positive half space in key order
    B.ncls = 0;
    for (int k = 0; k < 125; k++)
        B.key2cls[k] = -1;
    auto add_class = [&](int dx, int dy, int dz) {
        const int key = (dx + 2) + 5 * (dy + 2) + 25 * (dz + 2);
        B.key2cls[key] = B.ncls;
        B.cls_d[B.ncls][0] = dx;
        B.cls_d[B.ncls][1] = dy;
        B.cls_d[B.ncls][2] = dz;
        B.ncls++;
    };
    add_class(0, 0, 0);
    for (int dz = 0; dz <= 2; dz++)
        for (int dy = -2; dy <= 2; dy++)
            for (int dx = -2; dx <= 2; dx++) {
                const bool positive = dz > 0 || (dz == 0 && dy > 0) || (dz == 0 && dy == 0 && dx > 0);
                if (positive && flags[(dx + 2) + 5 * (dy + 2) + 25 * (dz + 2)]) {
                    if (B.ncls >= MAXCLS) {
                        cudaFree(ic);
                        lpmb_set_error("brick SpMV: more than %d displacement classes", MAXCLS);
                        return LPMB_ERR_UNSUPPORTED;
                    }
                    add_class(dx, dy, dz);
                }
            }
    LPMB_TRY(brick_upload_tables(B));
    // needed z-range of every (brick layer, class) tile: rows z with z < nz and (z owned or z + dz owned), dz >= 0
    B.h_zr.assign((size_t)B.nbz * B.ncls * 2, 0);
    const bool trim = param(c, "brick_trim", 1.0) != 0.0;
    for (int bz = 0; bz < B.nbz; bz++)
        for (int u = 0; u < B.ncls; u++) {
            const int need0 = trim ? std::max(0, B.own_z0 - B.cls_d[u][2]) : 0, need1 = trim ? std::min(B.nz, B.own_z1) : BE * B.nbz;
            const int z0 = std::min(BE, std::max(0, need0 - BE * bz)), z1 = std::min(BE, std::max(0, need1 - BE * bz));
            B.h_zr[((size_t)bz * B.ncls + u) * 2] = (unsigned char)(z1 > z0 ? z0 : 0);
            B.h_zr[((size_t)bz * B.ncls + u) * 2 + 1] = (unsigned char)(z1 > z0 ? z1 : 0);
        }
    LPMB_CUDA(cudaMalloc(&B.zr, B.h_zr.size()));
    LPMB_H2D(c, B.zr, B.h_zr.data(), B.h_zr.size());
    const size_t nent = (size_t)B.nbricks * B.ncls * BR;
    LPMB_CUDA(cudaMalloc(&B.bval, nent * 9 * sizeof(double)));
    LPMB_CUDA(cudaMalloc(&B.stage, (size_t)B.nbricks * 3 * NSLOT * sizeof(double)));
    LPMB_MEMSET(c, B.stage, 0, (size_t)B.nbricks * 3 * NSLOT * sizeof(double));
    for (double **v : {&B.ypart, &B.r, &B.p, &B.ap, &B.x, &B.b, &B.mask}) {
        LPMB_CUDA(cudaMalloc(v, (size_t)3 * B.P * sizeof(double)));
        LPMB_MEMSET(c, *v, 0, (size_t)3 * B.P * sizeof(double));
    }
    B.ic = ic;
    // 216.6 KB of dynamic shared memory per CTA: opt in on this context's device
    LPMB_CUDA(cudaFuncSetAttribute(brick_spmv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BrickSmem)));
    LPMB_CUDA(cudaFuncSetAttribute(brick_spmv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BrickSmem)));
    B.pattern_ready = true;
    return LPMB_OK;
}

static int brick_fill_values(lpmb_ctx *c, BrickMatrix &B)
{
    const size_t nent = (size_t)B.nbricks * B.ncls * BR;
    LPMB_CUDA(cudaMemsetAsync(B.bval, 0, nent * 9 * sizeof(double), c->stream));
    LPMB_TRY(brick_upload_tables(B));
    brick_fill_kernel<<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, c->Np, B.ic, B.inv, c->K.sptr, c->K.col,
                                                                     c->K.val, c->K.nbc, B.ncls, B.bval);
    LPMB_LAUNCH_CHECK(c);
    B.values_ready = true;
    return LPMB_OK;
}

extern "C" int lpmb_matrix_enable_bricks(lpmb_ctx *c, int on)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    if (!on) {
        lpmb_brick_release(c);
        return LPMB_OK;
    }
    BrickMatrix &B = g_bricks[c];
    if (!B.pattern_ready) {
        int rc = brick_build_pattern(c, B);
        if (c->world > 1) {
            // collective call: either every slab gets its bricks or none does (a rank that went on alone into the
            // peer set-up below would wait for the others forever)
            double *d_fail;
            LPMB_CUDA(cudaMalloc(&d_fail, sizeof(double)));
            const double mine = rc == LPMB_OK ? 0.0 : 1.0;
            double all = 0.0;
            LPMB_CUDA(cudaMemcpyAsync(d_fail, &mine, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            LPMB_TRY(lpmb_dist_allreduce_sum(c, d_fail, 1));
            LPMB_CUDA(cudaMemcpyAsync(&all, d_fail, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(d_fail);
            if (all != 0.0 && rc == LPMB_OK) {
                lpmb_set_error("brick SpMV: not eligible on %d other rank(s)", (int)all);
                rc = LPMB_ERR_UNSUPPORTED;
            }
        }
        if (rc != LPMB_OK) {
            lpmb_brick_release(c);
            return rc;
        }
    }
    B.enabled = true;
    // slab runs: let the neighbours push their boundary rows of the search direction straight into B.p (collective)
    if (c->world > 1)
        LPMB_TRY(lpmb_peer_halo_setup(c, B.p, B.P, B.inv));
    return LPMB_OK;
}

extern "C" long long lpmb_spmv_bytes_bricks(lpmb_ctx *c) { return c ? lpmb_brick_bytes(c) : 0; }

long long lpmb_brick_bytes(lpmb_ctx *c)
{
    auto it = g_bricks.find(c);
    if (it == g_bricks.end())
        return 0;
    const BrickMatrix &B = it->second;
    long long nent = 0;  // matrix entries actually streamed: the needed z-layers of every tile
    for (int bz = 0; bz < B.nbz; bz++)
        for (int u = 0; u < B.ncls; u++)
            nent += (long long)B.nbx * B.nby * 64 * (B.h_zr[((size_t)bz * B.ncls + u) * 2 + 1] - B.h_zr[((size_t)bz * B.ncls + u) * 2]);
    const long long halo = (long long)B.nbricks * (NSLOT - BR) * 3 * 8;
    return nent * 72 + 2 * halo + 3 * B.P * 8 * 4;  // matrix + staging (w+r) + x, ypart (w+r), y
}

// make sure the brick values mirror K.val
int lpmb_brick_prepare(lpmb_ctx *c)
{
    BrickMatrix &B = g_bricks[c];
    LPMB_REQUIRE(c->K.values_ready, LPMB_ERR_STATE, "stiffness matrix not available");
    if (!B.values_ready)
        LPMB_TRY(brick_fill_values(c, B));
    return LPMB_OK;
}

// CG work vectors in the permuted space
void lpmb_brick_vectors(lpmb_ctx *c, double **r, double **p, double **ap, double **x, double **b, double **mask, long long *P)
{
    BrickMatrix &B = g_bricks[c];
    *r = B.r;
    *p = B.p;
    *ap = B.ap;
    *x = B.x;
    *b = B.b;
    *mask = B.mask;
    *P = B.P;
}

int lpmb_brick_to_perm(lpmb_ctx *c, const double *src, double *dst)
{
    BrickMatrix &B = g_bricks[c];
    brick_to_perm_kernel<<<lpmb_blocks(3 * B.P, 256), 256, 0, c->stream>>>(B.P, c->Np, B.perm, src, dst, 0.0);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

int lpmb_brick_from_perm(lpmb_ctx *c, const double *src, double *dst)
{
    BrickMatrix &B = g_bricks[c];
    brick_from_perm_kernel<<<lpmb_blocks(3LL * c->N, 256), 256, 0, c->stream>>>(c->N, c->Np, B.P, B.inv, src, dst);
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

// slab runs: halo exchange of a brick-ordered vector, staged through the context's original-order CG vector
int lpmb_brick_exchange(lpmb_ctx *c, double *perm_vec)
{
    if (c->world <= 1)
        return LPMB_OK;
    BrickMatrix &B = g_bricks[c];
    const int own0 = lpmb_own0(c), own1 = lpmb_own1(c);
    auto range = [&](int i0, int count, int to_perm) -> int {
        if (count <= 0)
            return LPMB_OK;
        brick_range_kernel<<<lpmb_blocks(3LL * count, 256), 256, 0, c->stream>>>(i0, count, c->Np, B.P, B.inv, perm_vec, c->cg.p, to_perm);
        LPMB_LAUNCH_CHECK(c);
        return LPMB_OK;
    };
    if (c->rank > 0)
        LPMB_TRY(range(own0, c->narrow_send_lo, 0));
    if (c->rank < c->world - 1)
        LPMB_TRY(range(own1 - c->narrow_send_hi, c->narrow_send_hi, 0));
    LPMB_TRY(lpmb_dist_exchange(c, c->cg.p, 3, false));
    if (c->rank > 0)
        LPMB_TRY(range(own0 - c->narrow_recv_lo, c->narrow_recv_lo, 1));
    if (c->rank < c->world - 1)
        LPMB_TRY(range(own1, c->narrow_recv_hi, 1));
    return LPMB_OK;
}

// y = [mask .*] K x in the permuted space (+ p.Ap partials into `partials`, one per block of the gather grid)
int lpmb_brick_spmv(lpmb_ctx *c, const double *x, double *y, bool dot, const double *mask, double *partials, const double *scal, int gather_grid,
                    const PeerWait &halo_wait, const PeerPublish &pub)
{
    BrickMatrix &B = g_bricks[c];
    const int grid = B.nbricks < c->sm_count ? B.nbricks : c->sm_count;
    // experimental: wait for the halo pushes brick by brick (see the kernel); only meaningful when there is something to wait for
    const bool lazy = halo_wait.n == 2 && B.nbz > 1 && param(c, "brick_lazy_wait", 0.0) != 0.0;
    if (lazy)
        brick_spmv_kernel<true><<<grid, BR, sizeof(BrickSmem), c->stream>>>(B.nbricks, B.ncls, B.nbx, B.nby, B.nbz, B.P, B.bval, B.zr, x, B.ypart,
                                                                            B.stage, dot ? scal : nullptr, halo_wait, B.nbx * B.nby, B.own_z0,
                                                                            B.own_z1);
    else
        brick_spmv_kernel<false><<<grid, BR, sizeof(BrickSmem), c->stream>>>(B.nbricks, B.ncls, B.nbx, B.nby, B.nbz, B.P, B.bval, B.zr, x, B.ypart,
                                                                             B.stage, dot ? scal : nullptr, halo_wait, 0, 0, 0);
    LPMB_LAUNCH_CHECK(c);
    if (dot)
        brick_gather_kernel<true><<<gather_grid, 256, 0, c->stream>>>(B.P, B.nbx, B.nby, B.nbz, B.ypart, B.stage, mask, x, y, partials, scal, pub);
    else
        brick_gather_kernel<false><<<gather_grid, 256, 0, c->stream>>>(B.P, B.nbx, B.nby, B.nbz, B.ypart, B.stage, mask, x, y, nullptr, nullptr,
                                                                       PeerPublish());
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}
