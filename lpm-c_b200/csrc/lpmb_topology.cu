// lpmb_topology.cu -- O(N) neighbour search, AFEM connectivity and the uniform cell grid they share
// with the nonlocal-damage gather (compiled -fmad=false so the shell tests see the reference's
// distances bit for bit).
//
// Replaces, in the reference (both are O(N^2) all-pairs loops there):
//   searchNormalNeighbor()  src/neighbor.c:9-46    neighbors / nsign / nb / distance_initial / cs*_initial
//   searchAFEMNeighbor()    src/neighbor.c:49-141  conn / nb_conn / K_pointer (+ IK/JK/K_global sizing)
//
// Particles are binned into a uniform grid of cell size >= 1.01*cutoff2 (counting sort; the ids inside
// a cell are then sorted so everything downstream is deterministic).  A particle's candidates are the
// 27 surrounding cells; accepted neighbours are sorted by index, which reproduces the reference's
// "ascending j, shells interleaved" order (neighbor.c:16-41).  conn[i] is the sorted unique union of
// {j} U N1(j) over first-shell neighbours and {j} U N2(j) over second-shell neighbours
// (neighbor.c:56-112).
#include <cfloat>

#include "lpmb_internal.cuh"

static std::map<lpmb_ctx *, CellGrid> g_grids;  // one grid per context (set-up data, not on the timed path)

__global__ void minmax_kernel(const double *__restrict__ xyz, int N, int Np, double *__restrict__ out /* [grid][6] */)
{
    __shared__ double sm[6][256];
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256)
        for (int k = 0; k < 3; k++) {
            const double v = xyz[(size_t)k * Np + i];
            lo[k] = fmin(lo[k], v);
            hi[k] = fmax(hi[k], v);
        }
    for (int k = 0; k < 3; k++) {
        sm[k][threadIdx.x] = lo[k];
        sm[3 + k][threadIdx.x] = hi[k];
    }
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int k = 0; k < 3; k++) {
                sm[k][threadIdx.x] = fmin(sm[k][threadIdx.x], sm[k][threadIdx.x + s]);
                sm[3 + k][threadIdx.x] = fmax(sm[3 + k][threadIdx.x], sm[3 + k][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6)
        out[blockIdx.x * 6 + threadIdx.x] = sm[threadIdx.x][0];
}

__device__ __forceinline__ int cell_coord(double x, double o, double inv, int n)
{
    int c = (int)floor((x - o) * inv);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void cell_count_kernel(const double *__restrict__ xyz, int N, int Np, CellGrid g, int *__restrict__ cell_of, int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const double inv = 1.0 / g.cell;
    const int cx = cell_coord(xyz[i], g.ox, inv, g.nx), cy = cell_coord(xyz[(size_t)Np + i], g.oy, inv, g.ny),
              cz = cell_coord(xyz[(size_t)2 * Np + i], g.oz, inv, g.nz);
    const int cidx = cx + g.nx * (cy + g.ny * cz);
    cell_of[i] = cidx;
    atomicAdd(&count[cidx], 1);
}

__global__ void cell_fill_kernel(int N, const int *__restrict__ cell_of, const int *__restrict__ start, int *__restrict__ cursor, int *__restrict__ items)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const int cidx = cell_of[i];
    const int slot = atomicAdd(&cursor[cidx], 1);
    items[start[cidx] + slot] = i;
}

__global__ void cell_sort_kernel(long long ncells, const int *__restrict__ start, int *__restrict__ items)
{
    const long long cidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= ncells)
        return;
    const int a = start[cidx], b = start[cidx + 1];
    for (int i = a + 1; i < b; i++) {  // insertion sort (cells hold a handful of particles)
        const int v = items[i];
        int j = i - 1;
        while (j >= a && items[j] > v) {
            items[j + 1] = items[j];
            j--;
        }
        items[j + 1] = v;
    }
}

// Build (or rebuild) the grid over the given [3][Np] coordinate field.
int lpmb_grid_build(lpmb_ctx *c, const double *d_xyz, double cell_size, CellGrid **out)
{
    CellGrid &g = g_grids[c];
    cudaFree(g.start);
    cudaFree(g.items);
    g.start = g.items = nullptr;
    const int N = c->N;
    const int nb = 64;
    double *d_mm;
    LPMB_CUDA(cudaMalloc(&d_mm, nb * 6 * sizeof(double)));
    minmax_kernel<<<nb, 256, 0, c->stream>>>(d_xyz, N, c->Np, d_mm);
    LPMB_LAUNCH_CHECK(c);
    std::vector<double> mm(nb * 6);
    LPMB_CUDA(cudaMemcpyAsync(mm.data(), d_mm, nb * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_mm);
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int b = 0; b < nb; b++)
        for (int k = 0; k < 3; k++) {
            lo[k] = mm[b * 6 + k] < lo[k] ? mm[b * 6 + k] : lo[k];
            hi[k] = mm[b * 6 + 3 + k] > hi[k] ? mm[b * 6 + 3 + k] : hi[k];
        }
    g.cell = cell_size;
    g.ox = lo[0] - 0.5 * cell_size;
    g.oy = lo[1] - 0.5 * cell_size;
    g.oz = lo[2] - 0.5 * cell_size;
    g.nx = (int)floor((hi[0] - g.ox) / cell_size) + 1;
    g.ny = (int)floor((hi[1] - g.oy) / cell_size) + 1;
    g.nz = (int)floor((hi[2] - g.oz) / cell_size) + 1;
    g.ncells = (long long)g.nx * g.ny * g.nz;
    LPMB_REQUIRE(g.ncells > 0 && g.ncells < (1LL << 31) - 2, LPMB_ERR_UNSUPPORTED, "cell grid %d x %d x %d too large", g.nx, g.ny, g.nz);
    int *cell_of, *count;
    LPMB_CUDA(cudaMalloc(&cell_of, (size_t)N * sizeof(int)));
    LPMB_CUDA(cudaMalloc(&count, ((size_t)g.ncells + 1) * sizeof(int)));
    LPMB_CUDA(cudaMemsetAsync(count, 0, ((size_t)g.ncells + 1) * sizeof(int), c->stream));
    cell_count_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(d_xyz, N, c->Np, g, cell_of, count);
    LPMB_LAUNCH_CHECK(c);
    std::vector<int> h((size_t)g.ncells + 1);
    LPMB_CUDA(cudaMemcpyAsync(h.data(), count, ((size_t)g.ncells + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    int acc = 0;
    for (long long k = 0; k <= g.ncells; k++) {  // exclusive scan (set-up only)
        const int v = h[k];
        h[k] = acc;
        acc += v;
    }
    LPMB_CUDA(cudaMalloc(&g.start, ((size_t)g.ncells + 1) * sizeof(int)));
    LPMB_CUDA(cudaMalloc(&g.items, (size_t)N * sizeof(int)));
    LPMB_CUDA(cudaMemcpyAsync(g.start, h.data(), ((size_t)g.ncells + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    LPMB_CUDA(cudaMemsetAsync(count, 0, ((size_t)g.ncells + 1) * sizeof(int), c->stream));
    cell_fill_kernel<<<lpmb_blocks(N, 256), 256, 0, c->stream>>>(N, cell_of, g.start, count, g.items);
    LPMB_LAUNCH_CHECK(c);
    cell_sort_kernel<<<lpmb_blocks(g.ncells, 256), 256, 0, c->stream>>>(g.ncells, g.start, g.items);
    LPMB_LAUNCH_CHECK(c);
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(cell_of);
    cudaFree(count);
    if (out)
        *out = &g;
    return LPMB_OK;
}

void lpmb_grid_release(lpmb_ctx *c)
{
    auto it = g_grids.find(c);
    if (it != g_grids.end()) {
        cudaFree(it->second.start);
        cudaFree(it->second.items);
        g_grids.erase(it);
    }
}

// ---------------------------------------------------------------------------------------------
// searchNormalNeighbor   neighbor.c:9-46
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
neighbor_search_kernel(int N, int Np, int nn, CellGrid g, const double *__restrict__ xyz, double c1 /*1.01*cutoff1*/, double c2 /*1.01*cutoff2*/,
                       int *__restrict__ nbr, signed char *__restrict__ nsign, int *__restrict__ overflow)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= N)
        return;
    const double xi = xyz[i], yi = xyz[(size_t)Np + i], zi = xyz[(size_t)2 * Np + i];
    const double inv = 1.0 / g.cell;
    const int cx = cell_coord(xi, g.ox, inv, g.nx), cy = cell_coord(yi, g.oy, inv, g.ny), cz = cell_coord(zi, g.oz, inv, g.nz);
    int ids[32];
    signed char sh[32];
    int cnt = 0;
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
                    const double ax = xyz[j] - xi, ay = xyz[(size_t)Np + j] - yi, az = xyz[(size_t)2 * Np + j] - zi;
                    const double dis = sqrt(ax * ax + ay * ay + az * az);
                    int s = -1;
                    if ((dis < c1) && (j != i))
                        s = 0;
                    else if ((dis > c1) && (dis < c2))
                        s = 1;
                    if (s < 0)
                        continue;
                    if (cnt >= nn || cnt >= 32) {
                        atomicExch(overflow, 1);
                        continue;
                    }
                    // insert keeping ascending j
                    int k = cnt - 1;
                    while (k >= 0 && ids[k] > j) {
                        ids[k + 1] = ids[k];
                        sh[k + 1] = sh[k];
                        k--;
                    }
                    ids[k + 1] = j;
                    sh[k + 1] = (signed char)s;
                    cnt++;
                }
            }
        }
    }
    for (int k = 0; k < nn; k++) {
        nbr[(size_t)k * Np + i] = k < cnt ? ids[k] : -1;
        nsign[(size_t)k * Np + i] = k < cnt ? sh[k] : (signed char)-1;
    }
}

// ---------------------------------------------------------------------------------------------
// searchAFEMNeighbor   neighbor.c:49-112  -> conn[N][nconn] row-major (reference layout), -1 padded
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sorted_insert(int *arr, int &n, int cap, int v, int *overflow)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (arr[mid] == v)
            return;
        if (arr[mid] < v)
            lo = mid + 1;
        else
            hi = mid - 1;
    }
    if (n >= cap) {
        atomicExch(overflow, 2);
        return;
    }
    for (int k = n; k > lo; k--)
        arr[k] = arr[k - 1];
    arr[lo] = v;
    n++;
}

__global__ void __launch_bounds__(128)
afem_conn_kernel(int N, int Np, int nn, int nconn, const int *__restrict__ nbr, const signed char *__restrict__ nsign, int *__restrict__ conn,
                 int *__restrict__ overflow)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= N)
        return;
    int set[128];
    int n = 0;
    for (int j = 0; j < nn; j++) {
        const int nj = nbr[(size_t)j * Np + i];
        if (nj < 0)
            break;
        const int s = nsign[(size_t)j * Np + i];
        sorted_insert(set, n, nconn, nj, overflow);
        for (int m = 0; m < nn; m++) {
            const int nm = nbr[(size_t)m * Np + nj];
            if (nm < 0)
                break;
            if (nsign[(size_t)m * Np + nj] == s)
                sorted_insert(set, n, nconn, nm, overflow);
        }
    }
    for (int k = 0; k < nconn; k++)
        conn[(size_t)i * nconn + k] = k < n ? set[k] : -1;
}

extern "C" int lpmb_build_topology(lpmb_ctx *c, double cutoff1, double cutoff2)
{
    LPMB_REQUIRE(c && cutoff1 > 0 && cutoff2 > cutoff1, LPMB_ERR_ARG, "lpmb_build_topology: bad cutoffs");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_REQUIRE(c->fields.count("xyz"), LPMB_ERR_STATE, "xyz not uploaded");
    const double *xyz = fptr<double>(c, "xyz");
    CellGrid *g = nullptr;
    LPMB_TRY(lpmb_grid_build(c, xyz, 1.01 * cutoff2 * 1.0001, &g));
    int *nbr = fptr<int>(c, "neighbors");
    signed char *nsign = fptr<signed char>(c, "nsign");
    int *d_over;
    LPMB_CUDA(cudaMalloc(&d_over, sizeof(int)));
    LPMB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), c->stream));
    neighbor_search_kernel<<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, c->Np, c->nn, *g, xyz, 1.01 * cutoff1, 1.01 * cutoff2, nbr, nsign, d_over);
    LPMB_LAUNCH_CHECK(c);
    LPMB_TRY(lpmb_derive_topology(c, true));
    int over = 0;
    LPMB_CUDA(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_over);
    if (over) {
        lpmb_set_error("lpmb_build_topology: more than nneighbors entries for some particle (the reference would overrun its arrays)");
        return LPMB_ERR_ARG;
    }
    return lpmb_rebuild_connectivity(c);
}

// conn / nb_conn / K_pointer (and the block pattern of the stiffness matrix) from the neighbour lists on the device
// (neighbor.c:49-141); also used after a snapshot was loaded (lpmb_io.cu)
int lpmb_rebuild_connectivity(lpmb_ctx *c)
{
    int *nbr = fptr<int>(c, "neighbors");
    signed char *nsign = fptr<signed char>(c, "nsign");
    LPMB_REQUIRE(nbr && nsign, LPMB_ERR_STATE, "neighbour lists missing");
    int *d_over;
    LPMB_CUDA(cudaMalloc(&d_over, sizeof(int)));
    LPMB_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), c->stream));
    int *d_conn;
    LPMB_CUDA(cudaMalloc(&d_conn, (size_t)c->N * c->nconn * sizeof(int)));
    afem_conn_kernel<<<lpmb_blocks(c->N, 128), 128, 0, c->stream>>>(c->N, c->Np, c->nn, c->nconn, nbr, nsign, d_conn, d_over);
    LPMB_LAUNCH_CHECK(c);
    int over = 0;
    LPMB_CUDA(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_over);
    if (over) {
        cudaFree(d_conn);
        lpmb_set_error("lpmb_build_topology: more than nneighbors_AFEM+1 entries for some particle (the reference would overrun its arrays)");
        return LPMB_ERR_ARG;
    }
    const int rc = lpmb_set_connectivity_device(c, d_conn);
    cudaFree(d_conn);
    return rc;
}
