// lpmb_symspmv.cu -- EXPERIMENTAL symmetric-storage SpMV (off by default; parameter "spmv_symmetric" = 1).
//
// K is symmetric by construction (stiffness.c:441-481 stores one triangle).  The default SELL kernel
// (lpmb_solver.cu) streams both triangles: 61 blocks * 76 B per interior SC particle.  Here each row keeps
//   * its diagonal block and the blocks with column > row ("upper", SELL-32 like the full format), and
//   * for every lower neighbour j < i only two ints: j and the position of block (j,i) inside row j's upper
//     storage; the contribution K_ji^T x_j is gathered from there.
// Stored bytes per interior SC particle: 31*(72+4) + 31*8 + 48 = 2.65 KB instead of 4.69 KB.  Every stored block
// is touched twice -- by its own row and by the partner row -- so the saving only materialises if the second
// touch hits L2.  The loop is therefore interleaved (step t = upper block t + the lower entry whose partner
// keeps its block at upper index t), so that both touches are issued at the same step by the two warps.
//
// MEASURED (B200, real FD tangents, round 1): 1.11x faster than the full format at 48^3 and 100^3, but 0.85x at
// 216^3 (8.17 ms vs 6.97 ms): partners are up to 2*216^2 rows = 2 900 slices apart, ~9 500 warps are in flight
// and drift apart, the in-flight working set (~700 MB) is far larger than the 126 MB L2, so most second touches
// go back to HBM, and the partner gathers straddle two slices (3 lines per 256 B instead of 2).  A y-strip slice
// schedule ("sym_strip") did not change that.  Conclusion: exploiting symmetry on this chip needs the second
// use to come from the SAME SM (brick-blocked rows with shared-memory accumulation of the transposed products),
// not from L2 -- see DESIGN.md "what comes next".  The kernel is kept as a tested, optional variant.
#include <algorithm>

#include "lpmb_internal.cuh"

#define SYM_THREADS 128

void lpmb_sym_release(lpmb_ctx *c)
{
    SymMatrix &S = c->sym;
    cudaFree(S.usptr); cudaFree(S.lsptr); cudaFree(S.ucol); cudaFree(S.uval); cudaFree(S.lcol); cudaFree(S.lpos); cudaFree(S.order);
    S = SymMatrix();
}

__device__ __forceinline__ int sym_find(const int *__restrict__ col, const long long *__restrict__ sptr, const int *__restrict__ nbc, int row,
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

template <int D>
__global__ void sym_fill_kernel(int N, int nslices, const long long *__restrict__ sptr, const int *__restrict__ col, const double *__restrict__ val,
                                const int *__restrict__ nbc, const int *__restrict__ k0, const long long *__restrict__ usptr,
                                const long long *__restrict__ lsptr, long long ukunits, int *__restrict__ ucol, double *__restrict__ uval,
                                int *__restrict__ lcol, int *__restrict__ lpos)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = row >> 5, lane = row & 31;
    if (s >= nslices)
        return;
    const int n = row < N ? nbc[row] : 0, K0 = row < N ? k0[row] : 0;
    const int first_ge = n - K0;
    const long long ka = sptr[s];
    const int self = row < N ? row : 0;
    // upper part (diagonal block first: conn is sorted, the self block is the first with column >= row)
    const long long ua = usptr[s], ub = usptr[s + 1];
    for (long long u = ua; u < ub; u++) {
        const int m = (int)(u - ua);
        const bool live = m < K0;
        const long long k = ka + first_ge + m;
        ucol[u * 32 + lane] = live ? col[k * 32 + lane] : self;
#pragma unroll
        for (int e = 0; e < D * D; e++)
            uval[(u * D * D + e) * 32 + lane] = live ? val[(k * D * D + e) * 32 + lane] : 0.0;
    }
    // lower part: where does the partner keep block (j, row)?  Entry for partner j goes to lower slot
    // mj = index of that block in row j's upper storage, so that the owner's touch (its upper step mj) and
    // this row's touch (its lower step mj) happen at the same step of the interleaved loop -> one HBM fetch.
    const long long la = lsptr[s], lb = lsptr[s + 1];
    const int W = (int)(lb - la);
    for (long long l = la; l < lb; l++) {
        lcol[l * 32 + lane] = self;
        lpos[l * 32 + lane] = (int)(ukunits * 32 + lane);  // the all-zero padding unit
    }
    unsigned long long used = 0ull;
    for (int m = 0; m < first_ge; m++) {
        const int j = col[(ka + m) * 32 + lane];
        const int p = sym_find(col, sptr, nbc, j, row);
        if (p < 0)
            continue;  // asymmetric pattern entry: cannot happen for conn from neighbor.c
        const int mj = p - (nbc[j] - k0[j]);
        int slot = mj < W ? mj : W - 1;
        for (int tries = 0; tries < W && ((used >> slot) & 1ull); tries++)  // collisions only on ragged boundaries
            slot = (slot + 1) % W;
        used |= 1ull << slot;
        lcol[(la + slot) * 32 + lane] = j;
        lpos[(la + slot) * 32 + lane] = (int)((usptr[j >> 5] + mj) * 32 + (j & 31));
    }
}

__global__ void sym_zero_unit_kernel(double *__restrict__ uval, int *__restrict__ ucol, long long ukunits, int DD)
{
    const int t = threadIdx.x;
    if (t < 32)
        ucol[ukunits * 32 + t] = 0;
    for (int e = t; e < DD * 32; e += blockDim.x)
        uval[ukunits * DD * 32 + e] = 0.0;
}

__global__ void sym_slice_key_kernel(int s_begin, int s_end, int N, int Np, const double *__restrict__ x0, double *__restrict__ keys)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = s_begin + t;
    if (s >= s_end)
        return;
    int row = s * 32;
    if (row >= N)
        row = N - 1;
    keys[3 * t] = x0[row];
    keys[3 * t + 1] = x0[(size_t)Np + row];
    keys[3 * t + 2] = x0[(size_t)2 * Np + row];
}

int lpmb_sym_build(lpmb_ctx *c)
{
    SellMatrix &K = c->K;
    LPMB_REQUIRE(K.values_ready, LPMB_ERR_STATE, "stiffness matrix not available");
    SymMatrix &S = c->sym;
    const int N = c->N, Np = c->Np, D = c->dim, ns = K.nslices;
    if (!S.usptr) {
        // pattern-dependent part: built once per connectivity
        std::vector<int> h_nbc(Np), h_k0(Np);
        LPMB_CUDA(cudaMemcpy(h_nbc.data(), K.nbc, (size_t)Np * sizeof(int), cudaMemcpyDeviceToHost));
        LPMB_CUDA(cudaMemcpy(h_k0.data(), K.k0, (size_t)Np * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<long long> hu(ns + 1), hl(ns + 1);
        long long au = 0, al = 0;
        int wmax = 1;
        for (int r = 0; r < Np; r++)
            wmax = std::max(wmax, h_k0[r]);
        LPMB_REQUIRE(wmax <= 64, LPMB_ERR_UNSUPPORTED, "symmetric format: %d upper blocks in a row (max 64)", wmax);
        for (int s = 0; s < ns; s++) {
            hu[s] = au;
            hl[s] = al;
            int wu = 0, wl = 0;
            for (int l = 0; l < 32; l++) {
                const int r = s * 32 + l;
                wu = std::max(wu, h_k0[r]);
                wl = std::max(wl, h_nbc[r] - h_k0[r]);
            }
            au += wu;
            al += wl > 0 ? wmax : 0;  // lower slots are addressed by the partner's upper index (0..wmax-1)
        }
        hu[ns] = au;
        hl[ns] = al;
        S.ukunits = au;
        S.lkunits = al;
        LPMB_REQUIRE((au + 1) * 32 < (1LL << 31), LPMB_ERR_UNSUPPORTED, "symmetric format: %lld upper units exceed 32-bit positions", au);
        LPMB_CUDA(cudaMalloc(&S.usptr, (size_t)(ns + 1) * sizeof(long long)));
        LPMB_CUDA(cudaMalloc(&S.lsptr, (size_t)(ns + 1) * sizeof(long long)));
        LPMB_CUDA(cudaMemcpy(S.usptr, hu.data(), (size_t)(ns + 1) * sizeof(long long), cudaMemcpyHostToDevice));
        LPMB_CUDA(cudaMemcpy(S.lsptr, hl.data(), (size_t)(ns + 1) * sizeof(long long), cudaMemcpyHostToDevice));
        LPMB_CUDA(cudaMalloc(&S.ucol, (size_t)(au + 1) * 32 * sizeof(int)));
        LPMB_CUDA(cudaMalloc(&S.uval, (size_t)(au + 1) * D * D * 32 * sizeof(double)));
        LPMB_CUDA(cudaMalloc(&S.lcol, (size_t)std::max(al, 1LL) * 32 * sizeof(int)));
        LPMB_CUDA(cudaMalloc(&S.lpos, (size_t)std::max(al, 1LL) * 32 * sizeof(int)));
        // processing order of the slices that hold owned rows
        const int sb = lpmb_own0(c) / 32, se = (lpmb_own1(c) + 31) / 32;
        S.norder = se - sb;
        std::vector<int> order(S.norder);
        for (int t = 0; t < S.norder; t++)
            order[t] = sb + t;
        const double strip_rows = param(c, "sym_strip", 0.0);
        if (strip_rows > 0 && c->params.count("radius") && c->fields.count("xyz_initial")) {
            double *d_keys;
            LPMB_CUDA(cudaMalloc(&d_keys, (size_t)S.norder * 3 * sizeof(double)));
            sym_slice_key_kernel<<<lpmb_blocks(S.norder, 256), 256, 0, c->stream>>>(sb, se, N, Np, fptr<double>(c, "xyz_initial"), d_keys);
            LPMB_LAUNCH_CHECK(c);
            std::vector<double> keys((size_t)S.norder * 3);
            LPMB_CUDA(cudaMemcpyAsync(keys.data(), d_keys, keys.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(d_keys);
            const double w = strip_rows * 2.0 * param(c, "radius");
            double ymin = keys[1];
            for (int t = 0; t < S.norder; t++)
                ymin = std::min(ymin, keys[3 * t + 1]);
            std::vector<long long> strip(S.norder);
            for (int t = 0; t < S.norder; t++)
                strip[t] = (long long)floor((keys[3 * t + 1] - ymin) / w);
            std::vector<int> idx(S.norder);
            for (int t = 0; t < S.norder; t++)
                idx[t] = t;
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
                if (strip[a] != strip[b])
                    return strip[a] < strip[b];
                if (keys[3 * a + 2] != keys[3 * b + 2])
                    return keys[3 * a + 2] < keys[3 * b + 2];  // z
                return a < b;                                     // then the original (y, x) order
            });
            for (int t = 0; t < S.norder; t++)
                order[t] = sb + idx[t];
        }
        LPMB_CUDA(cudaMalloc(&S.order, (size_t)std::max(S.norder, 1) * sizeof(int)));
        LPMB_CUDA(cudaMemcpy(S.order, order.data(), (size_t)S.norder * sizeof(int), cudaMemcpyHostToDevice));
    }
    sym_zero_unit_kernel<<<1, 256, 0, c->stream>>>(S.uval, S.ucol, S.ukunits, D * D);
    LPMB_LAUNCH_CHECK(c);
    const int blocks = lpmb_blocks(Np, 128);
    if (D == 3)
        sym_fill_kernel<3><<<blocks, 128, 0, c->stream>>>(N, ns, K.sptr, K.col, K.val, K.nbc, K.k0, S.usptr, S.lsptr, S.ukunits, S.ucol, S.uval, S.lcol, S.lpos);
    else
        sym_fill_kernel<2><<<blocks, 128, 0, c->stream>>>(N, ns, K.sptr, K.col, K.val, K.nbc, K.k0, S.usptr, S.lsptr, S.ukunits, S.ucol, S.uval, S.lcol, S.lpos);
    LPMB_LAUNCH_CHECK(c);
    S.ready = true;
    return LPMB_OK;
}

long long lpmb_sym_bytes(lpmb_ctx *c)
{
    const SymMatrix &S = c->sym;
    const long long d = c->dim;
    return S.ukunits * 32 * (8 * d * d + 4) + S.lkunits * 32 * 8 + 16LL * (c->K.nslices + 1) + 16LL * d * c->Np;
}

__device__ __forceinline__ double sym_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = [mask .*] (K x) from the upper blocks; optional partial dot products x.y per block (deterministic)
template <int D, bool DOT>
__global__ void __launch_bounds__(SYM_THREADS)
sym_spmv_kernel(const int *__restrict__ order, int norder, const long long *__restrict__ usptr, const int *__restrict__ ucol,
                const double *__restrict__ uval, const long long *__restrict__ lsptr, const int *__restrict__ lcol, const int *__restrict__ lpos,
                const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ mask, int Np, double *__restrict__ partials,
                const double *__restrict__ scal)
{
    __shared__ double red[SYM_THREADS / 32];
    if (DOT && scal && scal[7] != 0.0)  // S_DONE
        return;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * SYM_THREADS + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * SYM_THREADS) >> 5;
    double dot = 0.0;
    for (int t = warp; t < norder; t += nwarps) {
        const int s = order[t];
        double acc[D];
#pragma unroll
        for (int r = 0; r < D; r++)
            acc[r] = 0.0;
        // Interleaved: step t handles upper block t of this slice AND the lower entry whose partner keeps its
        // block at upper index t, so both touches of a block are issued at the same step by the two warps.
        const long long ka = usptr[s], wu = usptr[s + 1] - ka;
        const long long la = lsptr[s], wl = lsptr[s + 1] - la;
        const long long W = wu > wl ? wu : wl;
        for (long long t = 0; t < W; t++) {
            if (t < wu) {
                const int cidx = __ldg(ucol + (ka + t) * 32 + lane);
                const double *vp = uval + (ka + t) * (D * D * 32) + lane;
                double a[D * D];
#pragma unroll
                for (int e = 0; e < D * D; e++)
                    a[e] = vp[e * 32];
                double xv[D];
#pragma unroll
                for (int q = 0; q < D; q++)
                    xv[q] = __ldg(x + (size_t)q * Np + cidx);
#pragma unroll
                for (int r = 0; r < D; r++)
#pragma unroll
                    for (int q = 0; q < D; q++)
                        acc[r] = fma(a[r * D + q], xv[q], acc[r]);
            }
            if (t < wl) {
                const int j = __ldcs(lcol + (la + t) * 32 + lane);
                const int pos = __ldcs(lpos + (la + t) * 32 + lane);
                const double *bp = uval + (size_t)(pos >> 5) * (D * D * 32) + (pos & 31);
                double a[D * D];
#pragma unroll
                for (int e = 0; e < D * D; e++)
                    a[e] = bp[e * 32];
                double xv[D];
#pragma unroll
                for (int q = 0; q < D; q++)
                    xv[q] = __ldg(x + (size_t)q * Np + j);
#pragma unroll
                for (int r = 0; r < D; r++)
#pragma unroll
                    for (int q = 0; q < D; q++)
                        acc[r] = fma(a[q * D + r], xv[q], acc[r]);
            }
        }
        const int row = s * 32 + lane;
#pragma unroll
        for (int r = 0; r < D; r++) {
            double v = acc[r];
            if (mask)
                v *= mask[(size_t)r * Np + row];
            y[(size_t)r * Np + row] = v;
            if (DOT)
                dot = fma(v, x[(size_t)r * Np + row], dot);
        }
    }
    if (DOT) {
        dot = sym_warp_sum(dot);
        const int w = threadIdx.x >> 5;
        if (lane == 0)
            red[w] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = 0.0;
#pragma unroll
            for (int i = 0; i < SYM_THREADS / 32; i++)
                tsum += red[i];
            partials[blockIdx.x] = tsum;
        }
    }
}

int lpmb_sym_spmv(lpmb_ctx *c, const double *x, double *y, bool dot, const double *mask, double *partials, const double *scal, int grid)
{
    SymMatrix &S = c->sym;
    if (c->dim == 3) {
        if (dot)
            sym_spmv_kernel<3, true><<<grid, SYM_THREADS, 0, c->stream>>>(S.order, S.norder, S.usptr, S.ucol, S.uval, S.lsptr, S.lcol, S.lpos, x, y, mask, c->Np, partials, scal);
        else
            sym_spmv_kernel<3, false><<<grid, SYM_THREADS, 0, c->stream>>>(S.order, S.norder, S.usptr, S.ucol, S.uval, S.lsptr, S.lcol, S.lpos, x, y, mask, c->Np, nullptr, nullptr);
    } else {
        if (dot)
            sym_spmv_kernel<2, true><<<grid, SYM_THREADS, 0, c->stream>>>(S.order, S.norder, S.usptr, S.ucol, S.uval, S.lsptr, S.lcol, S.lpos, x, y, mask, c->Np, partials, scal);
        else
            sym_spmv_kernel<2, false><<<grid, SYM_THREADS, 0, c->stream>>>(S.order, S.norder, S.usptr, S.ucol, S.uval, S.lsptr, S.lcol, S.lpos, x, y, mask, c->Np, nullptr, nullptr);
    }
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}
