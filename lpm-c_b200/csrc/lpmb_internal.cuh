// lpmb_internal.cuh -- context, field registry and launch helpers shared by all translation units.
// Not part of the C ABI (include/lpmb200.h is).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "lpmb200.h"

#define LPMB_SLICE 32  // SELL slice height = warp size; also the padding quantum of every SoA array

void lpmb_set_error(const char *fmt, ...);

#define LPMB_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            lpmb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return LPMB_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)

// Every device operation of a context is ordered on c->stream, which is created cudaStreamNonBlocking: the legacy
// NULL stream does NOT synchronise with it, so memsets and copies must be issued on c->stream as well (a cudaMemset /
// cudaMemcpy on the NULL stream may run before or after kernels queued on c->stream).  The copies wait for completion
// because their host side is pageable and reused right away.
#define LPMB_MEMSET(c, p, v, n) LPMB_CUDA(cudaMemsetAsync((p), (v), (n), (c)->stream))
#define LPMB_D2H(c, dst, src, n)                                                                     \
    do {                                                                                             \
        LPMB_CUDA(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyDeviceToHost, (c)->stream));          \
        LPMB_CUDA(cudaStreamSynchronize((c)->stream));                                               \
    } while (0)
#define LPMB_H2D(c, dst, src, n)                                                                     \
    do {                                                                                             \
        LPMB_CUDA(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyHostToDevice, (c)->stream));          \
        LPMB_CUDA(cudaStreamSynchronize((c)->stream));                                               \
    } while (0)

#define LPMB_TRY(call)              \
    do {                            \
        int r__ = (call);           \
        if (r__ != LPMB_OK)         \
            return r__;             \
    } while (0)

#define LPMB_REQUIRE(cond, code, ...)  \
    do {                               \
        if (!(cond)) {                 \
            lpmb_set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

// ---- field registry ---------------------------------------------------------------------------
// Device layouts (Np = N rounded up to a multiple of 32):
//   FK_BOND   [nn][Np]     host [N][nn]        slot-major: thread i reads slot j at j*Np+i (coalesced)
//   FK_PART   [c][Np]      host [N][c]         component-major
//   FK_DOF    [dim][Np]    host [N*dim]        component-major copy of an interleaved DoF vector
//   FK_PIN    [3][Np]      host [N*3]          Pin keeps NDIM=3 stride even in 2-D (stiffness.c:527)
//   FK_RAW    [count]      host [count]        no re-layout
enum FieldKind { FK_BOND = 0, FK_PART = 1, FK_DOF = 2, FK_PIN = 3, FK_RAW = 4 };
enum FieldType { FT_F64 = 0, FT_I32 = 1, FT_I8 = 2 };

struct Field {
    void *d = nullptr;
    FieldKind kind = FK_RAW;
    FieldType type = FT_F64;
    int comps = 1;       // rows of the device layout (nn for bonds, c for parts, dim for dofs)
    size_t count = 0;    // device elements
    size_t elem() const { return type == FT_F64 ? 8 : (type == FT_I32 ? 4 : 1); }
};

struct SellMatrix {
    int D = 3;                 // block dimension (dim)
    int nslices = 0;
    long long kunits = 0;      // sum of slice widths
    long long nblocks = 0;     // sum nb_conn (algorithmic block count)
    long long nnz_upper = 0;   // K_pointer[N][1]
    long long *sptr = nullptr; // [nslices+1] offsets in k-units
    int *col = nullptr;        // [kunits][32] block column (pad: own row, value 0)
    double *val = nullptr;     // [kunits][D*D][32]
    int *nbc = nullptr;        // [Np] nb_conn
    int *k0 = nullptr;         // [Np] K_pointer[i][0]
    long long *kp = nullptr;   // [N+1] K_pointer[i][1] (64-bit)
    bool pattern_ready = false;
    bool values_ready = false;
    // Row-packed mirror for SMALL lattices (lpmb_solver.cu, spmv_rows_kernel): block row i = nb_conn[i] consecutive
    // D x D blocks, so one warp owns a row and its lanes stride over the blocks -- ~60x more parallelism per row than the
    // SELL kernel's one-lane-per-row walk, which is latency-bound when the whole lattice is a few hundred warps.
    long long *rptr = nullptr; // [N+1] block offsets
    int *rcol = nullptr;       // [nblocks]
    double *rval = nullptr;    // [nblocks][D*D]
    bool rows_ready = false;   // mirrors val (reset whenever val changes)
};

struct CGWork {
    double *r = nullptr, *p = nullptr, *ap = nullptr, *x = nullptr;  // [D][Np] each
    double *partials = nullptr;                                       // [2][max_blocks]
    double *scal = nullptr;                                           // device scalars (see lpmb_solver.cu)
    double *h_scal = nullptr;                                         // pinned mirror
    unsigned int *counters = nullptr;                                 // [4] last-block tickets (gather, update, direction)
    double *z = nullptr;                                              // [D][Np] preconditioned residual (fast mode only)
    int max_blocks = 0;
    // small lattices (row-packed SpMV, one GPU): a batch of 16 CG iterations = 48 launches replayed as ONE CUDA graph
    cudaGraphExec_t graph = nullptr;
    unsigned long long graph_key[12] = {0};
};

struct lpmb_ctx {
    int device = 0;
    int N = 0, Np = 0, dim = 3, lattice = 2, nn = 0, nconn = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    std::map<std::string, Field> fields;
    std::map<std::string, double> params;
    SellMatrix K;
    CGWork cg;
    double *mask = nullptr;      // [dim][Np] 1.0 free / 0.0 constrained (nullptr = all free)
    double *fd_tab = nullptr;    // FD assembly scratch: perturbed shell sums of a window of particles (lpmb_stiffness.cu)
    size_t fd_tab_bytes = 0;
    void *staging = nullptr;     // device staging for host<->device re-layout
    size_t staging_bytes = 0;
    void *h_staging = nullptr;   // pinned host staging
    size_t h_staging_bytes = 0;
    // multi-GPU
    void *nccl = nullptr;
    int rank = 0, world = 1;
    // slab decomposition: local particles are [ghost_lo | owned | ghost_hi] in global index order;
    // owned = [own0, own1).  narrow = ghost particles next to the owned range that the CG exchanges
    // every iteration (conn reach: 2 lattice layers); wide = all ghosts (4 layers: positions / state).
    int own0 = 0, own1 = -1;
    int narrow_recv_lo = 0, narrow_recv_hi = 0, narrow_send_lo = 0, narrow_send_hi = 0;
    int wide_send_lo = 0, wide_send_hi = 0;
    // optional live profiling of the dominant kernel (CUDA events around every CG SpMV launch)
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;
    double prof_spmv_ms = 0.0;
    long long prof_spmv_calls = 0;
};

// registry helpers (lpmb_ctx.cu)
Field *lpmb_field(lpmb_ctx *c, const char *name);
int lpmb_field_alloc(lpmb_ctx *c, const char *name, FieldKind kind, FieldType type, int comps);
int lpmb_ensure_staging(lpmb_ctx *c, size_t bytes);
int lpmb_ensure_h_staging(lpmb_ctx *c, size_t bytes);
template <typename T>
static inline T *fptr(lpmb_ctx *c, const char *name)
{
    Field *f = lpmb_field(c, name);
    return f ? reinterpret_cast<T *>(f->d) : nullptr;
}
static inline double param(lpmb_ctx *c, const char *name, double dflt = 0.0)
{
    auto it = c->params.find(name);
    return it == c->params.end() ? dflt : it->second;
}

static inline int lpmb_blocks(long long work, int threads) { return (int)((work + threads - 1) / threads); }

#define LPMB_LAUNCH_CHECK(c)                                                                   \
    do {                                                                                       \
        (c)->launches++;                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            lpmb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return LPMB_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

int lpmb_upload_soa_f64(lpmb_ctx *c, const double *host, double *d_dst, int comps);
int lpmb_download_soa_f64(lpmb_ctx *c, const double *d_src, double *host, int comps);

// uniform cell grid (lpmb_topology.cu), shared by the neighbour search and the nonlocal damage gather
struct CellGrid {
    double ox = 0, oy = 0, oz = 0, cell = 1;
    int nx = 1, ny = 1, nz = 1;
    int *start = nullptr;  // [ncells+1]
    int *items = nullptr;  // [N] particle ids, ascending inside each cell
    long long ncells = 0;
};
int lpmb_grid_build(lpmb_ctx *c, const double *d_xyz, double cell_size, CellGrid **out);
void lpmb_grid_release(lpmb_ctx *c);
int lpmb_set_connectivity_device(lpmb_ctx *c, const int *d_conn);
int lpmb_rebuild_connectivity(lpmb_ctx *c);   // conn / block pattern from the device neighbour lists (lpmb_topology.cu)
int lpmb_derive_topology(lpmb_ctx *c, bool initial_geometry);
int lpmb_compute_stress(lpmb_ctx *c);
int lpmb_refresh_mask(lpmb_ctx *c);
int lpmb_cp_return_map(lpmb_ctx *c);
// the same kernel on explicit input / output arrays (per-particle entry point: scratch twins); err_particle != nullptr:
// a singular slip Jacobian is reported there (particle index + 1, 0 = none) instead of failing the call
struct CPIO {
    const double *dL, *dLt, *TdLt, *csx, *csy, *csz;
    double *dLp2, *gy2, *A2, *As2, *ddLp, *RSS;
    int *Jact;
    double *dgy, *dA, *dAs;
    int *pl_flag;
};
int lpmb_cp_return_map_io(lpmb_ctx *c, const CPIO *io, int *err_particle);

// multi-GPU helpers (lpmb_dist.cu); all are no-ops when world == 1
static inline int lpmb_own0(const lpmb_ctx *c) { return c->own0; }
static inline int lpmb_own1(const lpmb_ctx *c) { return c->own1 < 0 ? c->N : c->own1; }
int lpmb_dist_exchange(lpmb_ctx *c, double *v, int comps, bool wide);   // halo exchange of a [comps][Np] array
int lpmb_dist_allreduce_sum(lpmb_ctx *c, double *d_buf, int count);     // in-stream, in place
void lpmb_dist_release(lpmb_ctx *c);

int lpmb_dist_allgather_bytes(lpmb_ctx *c, const void *d_send, void *d_recv, size_t bytes_per_rank);
int lpmb_dist_neighbor_doubles(lpmb_ctx *c, double *base, long long stride, int comps, long long send_lo_off, long long recv_lo_off,
                               long long send_hi_off, long long recv_hi_off, long long n);
int lpmb_dist_allgatherv_doubles(lpmb_ctx *c, double *base, long long stride, int comps, const long long *offs, const long long *counts);
int lpmb_dist_neighbor_ints(lpmb_ctx *c, const int *to_lo, int n_to_lo, const int *to_hi, int n_to_hi, int *from_lo, int n_from_lo, int *from_hi,
                            int n_from_hi);

// NVLink peer-memory fast path (lpmb_peer.cu): CUDA-IPC mapped buffers of the other ranks on the same box.
//  * scalar all-reduce: every rank stores its partial + a sequence number into every rank's slot array; the consuming
//    kernel spins (acquire, system scope) until all `world` slots carry the sequence, then sums them in rank order
//  * halo push (brick-ordered CG vectors): boundary rows are written straight into the neighbour's vector, a flag
//    follows; the neighbour's SpMV kernel waits for the flag
#define LPMB_PEER_MAXW 16
struct PeerWait {              // passed by value to consumer kernels; n == 0 -> nothing to wait for
    const unsigned long long *seqs = nullptr;
    unsigned long long seq = 0;
    int n = 0;
};
int lpmb_peer_init(lpmb_ctx *c);
void lpmb_peer_release(lpmb_ctx *c);
bool lpmb_peer_ready(lpmb_ctx *c);
// reduce `nparts` per-block partials to this rank's scalar, publish it to all ranks; *vals (world doubles, local
// memory) and *wait describe what the consumer kernel has to do.  Skipped on the device when scal[S_DONE] != 0.
int lpmb_peer_allreduce_publish(lpmb_ctx *c, const double *partials, int nparts, const double *scal, const double **vals, PeerWait *wait);
// The same without the extra launch: the PRODUCER kernel of the partials gets `pub` by value and its last block to
// finish folds the partials (same fixed order) and publishes (lpmb_last_block / lpmb_peer_publish_block below).
struct PeerPublish {
    double *vals[LPMB_PEER_MAXW];
    unsigned long long *seqs[LPMB_PEER_MAXW];
    unsigned long long seq = 0;
    unsigned int *counter = nullptr;  // last-block detection, self-resetting
    int world = 0, rank = 0, set = 0; // world == 0: nothing to publish
};
int lpmb_peer_allreduce_prepare(lpmb_ctx *c, unsigned int *counter, PeerPublish *pub, const double **vals, PeerWait *wait);
int lpmb_peer_halo_setup(lpmb_ctx *c, double *perm_vec, long long P, const int *inv);
bool lpmb_peer_halo_ready(lpmb_ctx *c);
int lpmb_peer_halo_push(lpmb_ctx *c, const double *scal, PeerWait *wait);
void lpmb_peer_halo_release(lpmb_ctx *c);

__device__ __forceinline__ unsigned long long lpmb_ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lpmb_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// True in every thread of the LAST block of the grid to arrive here; every write the other blocks made before their own
// call is then visible to it (threadfence + atomic ticket).  The counter resets itself for the next launch.  All threads
// of all blocks must call it.
__device__ __forceinline__ bool lpmb_last_block(unsigned int *counter)
{
    __shared__ int lpmb_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        lpmb_is_last = ticket == gridDim.x - 1;
        if (lpmb_is_last)
            *counter = 0;
    }
    __syncthreads();
    const bool last = lpmb_is_last != 0;
    if (last)
        __threadfence();
    return last;
}
// One block of 256 threads: fold `nparts` per-block partials in the fixed order of peer_publish_kernel (lpmb_peer.cu) and
// store {value, sequence} into slot [set][rank] of every rank's buffer.  Partials are read with ld.cg (written by other
// blocks of the same launch).
__device__ __forceinline__ void lpmb_peer_publish_block(const double *partials, int nparts, const PeerPublish &pb, double *red /* [8] */)
{
    __shared__ double lpmb_pub_total;
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256)
        s += __ldcg(partials + i);
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++)
            t += red[k];
        lpmb_pub_total = t;
    }
    __syncthreads();
    if ((int)threadIdx.x < pb.world) {
        const int r = threadIdx.x;
        *(volatile double *)(pb.vals[r] + pb.set * LPMB_PEER_MAXW + pb.rank) = lpmb_pub_total;
        __threadfence_system();
        lpmb_st_release_sys(pb.seqs[r] + pb.set * LPMB_PEER_MAXW + pb.rank, pb.seq);
    }
}
// thread 0 of the block waits for the peers, then the block proceeds
__device__ __forceinline__ void lpmb_peer_wait(const PeerWait &w)
{
    if (w.n > 0) {
        if (threadIdx.x == 0)
            for (int r = 0; r < w.n; r++)
                while (lpmb_ld_acquire_sys(w.seqs + r) < w.seq) {
                }
        __syncthreads();
    }
}


// brick-blocked symmetric SpMV (lpmb_brick.cu): optional, single GPU, simple-cubic 3-D
void lpmb_brick_release(lpmb_ctx *c);
void lpmb_brick_touch(lpmb_ctx *c);   // K.val changed
bool lpmb_brick_active(lpmb_ctx *c);
int lpmb_brick_prepare(lpmb_ctx *c);
void lpmb_brick_vectors(lpmb_ctx *c, double **r, double **p, double **ap, double **x, double **b, double **mask, long long *P);
int lpmb_brick_to_perm(lpmb_ctx *c, const double *src, double *dst);
int lpmb_brick_from_perm(lpmb_ctx *c, const double *src, double *dst);
struct PeerWait;
int lpmb_brick_spmv(lpmb_ctx *c, const double *x, double *y, bool dot, const double *mask, double *partials, const double *scal, int gather_grid,
                    const PeerWait &halo_wait, const PeerPublish &pub);
long long lpmb_brick_bytes(lpmb_ctx *c);
int lpmb_brick_exchange(lpmb_ctx *c, double *perm_vec);

// matrix-free multigrid preconditioner of the opt-in fast mode (lpmb_mg.cu)
int lpmb_mg_prepare(lpmb_ctx *c, const double *mask0);
int lpmb_mg_apply(lpmb_ctx *c, const double *r, double *z, const double *done);
void lpmb_mg_touch(lpmb_ctx *c);
void lpmb_mg_release(lpmb_ctx *c);
int lpmb_mg_levels(lpmb_ctx *c);

// solver-side entry points used across TUs
int lpmb_cg_alloc(lpmb_ctx *c);
int lpmb_matrix_alloc_values(lpmb_ctx *c);
