// lpmb_peer.cu -- NVLink peer-memory fast path for the traffic that happens every CG iteration.
//
// The reference has no distributed path at all; ours shards z-slabs over the GPUs of one
// box (lpmb_dist.cu).  A CG iteration needs one halo exchange of the search direction and two scalar all-reduces
// (solver.c:209-255 run on a slab).  Through NCCL each of the three costs a collective launch (~25-40 us), which
// at 8 GPUs is a quarter of the iteration.  Here the ranks map each other's buffers with CUDA IPC and talk with
// plain stores over NVLink:
//
//   all-reduce   producer kernel (one block): folds the per-block partials to the rank's scalar and stores
//                {value, sequence} into slot [seq&1][rank] of EVERY rank's slot array (value, __threadfence_system,
//                then the sequence with st.release.sys).  Consumer = the CG kernel that needs the sum: thread 0 of
//                each block spins (ld.acquire.sys, local memory) until all `world` slots carry the sequence, then
//                every block adds the `world` values in rank order -> the same bits on every rank, no NCCL call,
//                no extra launch on the consumer side.  Two slot sets alternate; a set is reused by all-reduce k+2,
//                which a rank can only reach after it consumed k+1, which needs every peer's k+1 producer, which
//                runs after that peer's consumer of k: no overwrite before use.
//   halo push    the boundary rows of the brick-ordered search direction are written straight into the
//                neighbour's vector (index tables exchanged once), the last block to finish raises a flag in the
//                neighbour's buffer; the neighbour's SpMV kernel waits for it before it reads x.  The write cannot
//                race with the neighbour's previous SpMV (the all-reduce of p.Ap in between orders them) nor with
//                its direction update (which leaves masked = ghost rows untouched in this mode).
//
// Everything is keyed by monotonically increasing sequence numbers kept identically on all ranks (they issue the
// same calls); kernels that early-out after convergence (scal[S_DONE]) skip both the publish and the wait on all
// ranks alike.  If IPC mapping is not possible on some rank, all ranks fall back to NCCL together.
#include "lpmb_internal.cuh"

#define PEER_BUF_BYTES 4096
#define OFF_VALS 0                                   // double [2][MAXW]
#define OFF_SEQS (2 * LPMB_PEER_MAXW * 8)            // u64    [2][MAXW]
#define OFF_HALO (4 * LPMB_PEER_MAXW * 8)            // u64    [2]: raised by rank-1 / rank+1
#define OFF_COUNTER (OFF_HALO + 64)                  // u32    [2]: last-block detection of the push kernel

struct PeerComm {
    bool ready = false;
    unsigned char *buf = nullptr;
    unsigned char *peer_buf[LPMB_PEER_MAXW] = {nullptr};
    unsigned long long seq = 0, halo_seq = 0;
    // halo push of one brick-ordered vector
    bool halo_ready = false;
    double *my_vec = nullptr;
    long long my_P = 0;
    double *nbr_vec[2] = {nullptr, nullptr};  // [0] = rank-1, [1] = rank+1 (IPC mappings)
    long long nbr_P[2] = {0, 0};
    int *src[2] = {nullptr, nullptr}, *dst[2] = {nullptr, nullptr};
    int cnt[2] = {0, 0};
};

static std::map<lpmb_ctx *, PeerComm> g_peers;

struct PeerTargets {
    double *vals[LPMB_PEER_MAXW];
    unsigned long long *seqs[LPMB_PEER_MAXW];
};

bool lpmb_peer_ready(lpmb_ctx *c)
{
    auto it = g_peers.find(c);
    return it != g_peers.end() && it->second.ready && param(c, "peer_comm", 1.0) != 0.0;
}

bool lpmb_peer_halo_ready(lpmb_ctx *c)
{
    auto it = g_peers.find(c);
    return lpmb_peer_ready(c) && it->second.halo_ready;
}

// all ranks agree on success (min over ranks) through NCCL
static int agree(lpmb_ctx *c, bool ok, bool *all_ok)
{
    double *d;
    LPMB_CUDA(cudaMalloc(&d, 8));
    const double v = ok ? 0.0 : 1.0;
    LPMB_CUDA(cudaMemcpyAsync(d, &v, 8, cudaMemcpyHostToDevice, c->stream));
    LPMB_TRY(lpmb_dist_allreduce_sum(c, d, 1));
    double s = 0.0;
    LPMB_CUDA(cudaMemcpyAsync(&s, d, 8, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    *all_ok = s == 0.0;
    return LPMB_OK;
}

// exchange one IPC handle per rank; open those in `want` (bitmask of ranks)
static int exchange_and_open(lpmb_ctx *c, void *local, unsigned want, void **opened /* [world] */, bool *ok)
{
    const int W = c->world;
    cudaIpcMemHandle_t mine;
    *ok = cudaIpcGetMemHandle(&mine, local) == cudaSuccess;
    if (!*ok)
        cudaGetLastError();
    unsigned char *d_send, *d_recv;
    LPMB_CUDA(cudaMalloc(&d_send, sizeof(mine)));
    LPMB_CUDA(cudaMalloc(&d_recv, sizeof(mine) * W));
    LPMB_CUDA(cudaMemcpyAsync(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    LPMB_TRY(lpmb_dist_allgather_bytes(c, d_send, d_recv, sizeof(mine)));
    std::vector<cudaIpcMemHandle_t> all(W);
    LPMB_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(mine) * W, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_send);
    cudaFree(d_recv);
    bool all_got;
    LPMB_TRY(agree(c, *ok, &all_got));  // nobody opens anything unless every handle is valid
    *ok = all_got;
    if (!all_got)
        return LPMB_OK;
    for (int r = 0; r < W; r++) {
        opened[r] = nullptr;
        if (r == c->rank) {
            opened[r] = local;
        } else if (want & (1u << r)) {
            if (cudaIpcOpenMemHandle(&opened[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                opened[r] = nullptr;
                *ok = false;
            }
        }
    }
    return LPMB_OK;
}

int lpmb_peer_init(lpmb_ctx *c)
{
    if (c->world <= 1 || c->world > LPMB_PEER_MAXW)
        return LPMB_OK;
    PeerComm &pc = g_peers[c];
    LPMB_CUDA(cudaMalloc(&pc.buf, PEER_BUF_BYTES));
    LPMB_MEMSET(c, pc.buf, 0, PEER_BUF_BYTES);
    {   // neighbours that do not exist never raise a flag: pre-raise it for good
        unsigned long long halo[2] = {c->rank > 0 ? 0ull : ~0ull, c->rank < c->world - 1 ? 0ull : ~0ull};
        LPMB_H2D(c, pc.buf + OFF_HALO, halo, sizeof(halo));
    }
    void *opened[LPMB_PEER_MAXW] = {nullptr};
    bool ok = false;
    LPMB_TRY(exchange_and_open(c, pc.buf, (1u << c->world) - 1u, opened, &ok));
    bool all_ok = false;
    LPMB_TRY(agree(c, ok, &all_ok));
    if (!all_ok) {
        for (int r = 0; r < c->world; r++)
            if (opened[r] && r != c->rank)
                cudaIpcCloseMemHandle(opened[r]);
        cudaFree(pc.buf);
        g_peers.erase(c);
        return LPMB_OK;  // NCCL path stays in use on every rank
    }
    for (int r = 0; r < c->world; r++)
        pc.peer_buf[r] = (unsigned char *)opened[r];
    pc.ready = true;
    return LPMB_OK;
}

void lpmb_peer_halo_release(lpmb_ctx *c)
{
    auto it = g_peers.find(c);
    if (it == g_peers.end())
        return;
    PeerComm &pc = it->second;
    for (int k = 0; k < 2; k++) {
        if (pc.nbr_vec[k])
            cudaIpcCloseMemHandle(pc.nbr_vec[k]);
        cudaFree(pc.src[k]);
        cudaFree(pc.dst[k]);
        pc.nbr_vec[k] = nullptr;
        pc.src[k] = pc.dst[k] = nullptr;
        pc.cnt[k] = 0;
    }
    pc.halo_ready = false;
}

void lpmb_peer_release(lpmb_ctx *c)
{
    auto it = g_peers.find(c);
    if (it == g_peers.end())
        return;
    lpmb_peer_halo_release(c);
    PeerComm &pc = it->second;
    for (int r = 0; r < c->world; r++)
        if (pc.peer_buf[r] && r != c->rank)
            cudaIpcCloseMemHandle(pc.peer_buf[r]);
    cudaFree(pc.buf);
    g_peers.erase(it);
}

// ---- scalar all-reduce ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_publish_kernel(const double *__restrict__ partials, int nparts, const double *__restrict__ scal, PeerTargets tg, int world, int rank, int set,
                    unsigned long long seq)
{
    __shared__ double red[8];
    __shared__ double total;
    if (scal && scal[7] != 0.0)  // S_DONE: every rank skips alike
        return;
    // fixed-order fold of the per-block partials (same scheme as reduce_partials in lpmb_solver.cu)
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256)
        s += partials[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++)
            t += red[k];
        total = t;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        *(volatile double *)(tg.vals[r] + set * LPMB_PEER_MAXW + rank) = total;
        __threadfence_system();
        lpmb_st_release_sys(tg.seqs[r] + set * LPMB_PEER_MAXW + rank, seq);
    }
}

int lpmb_peer_allreduce_publish(lpmb_ctx *c, const double *partials, int nparts, const double *scal, const double **vals, PeerWait *wait)
{
    PeerComm &pc = g_peers[c];
    pc.seq++;
    const int set = (int)(pc.seq & 1);
    PeerTargets tg;
    for (int r = 0; r < c->world; r++) {
        tg.vals[r] = reinterpret_cast<double *>(pc.peer_buf[r] + OFF_VALS);
        tg.seqs[r] = reinterpret_cast<unsigned long long *>(pc.peer_buf[r] + OFF_SEQS);
    }
    peer_publish_kernel<<<1, 256, 0, c->stream>>>(partials, nparts, scal, tg, c->world, c->rank, set, pc.seq);
    LPMB_LAUNCH_CHECK(c);
    *vals = reinterpret_cast<const double *>(pc.buf + OFF_VALS) + set * LPMB_PEER_MAXW;
    wait->seqs = reinterpret_cast<const unsigned long long *>(pc.buf + OFF_SEQS) + set * LPMB_PEER_MAXW;
    wait->seq = pc.seq;
    wait->n = c->world;
    return LPMB_OK;
}

// the same bookkeeping without a launch: the caller hands `pub` to the kernel that produces the partials
int lpmb_peer_allreduce_prepare(lpmb_ctx *c, unsigned int *counter, PeerPublish *pub, const double **vals, PeerWait *wait)
{
    PeerComm &pc = g_peers[c];
    pc.seq++;
    const int set = (int)(pc.seq & 1);
    for (int r = 0; r < c->world; r++) {
        pub->vals[r] = reinterpret_cast<double *>(pc.peer_buf[r] + OFF_VALS);
        pub->seqs[r] = reinterpret_cast<unsigned long long *>(pc.peer_buf[r] + OFF_SEQS);
    }
    pub->seq = pc.seq;
    pub->counter = counter;
    pub->world = c->world;
    pub->rank = c->rank;
    pub->set = set;
    *vals = reinterpret_cast<const double *>(pc.buf + OFF_VALS) + set * LPMB_PEER_MAXW;
    wait->seqs = reinterpret_cast<const unsigned long long *>(pc.buf + OFF_SEQS) + set * LPMB_PEER_MAXW;
    wait->seq = pc.seq;
    wait->n = c->world;
    return LPMB_OK;
}

// ---- halo push --------------------------------------------------------------------------------------
// my rows src[t] -> the neighbour's rows dst[t], three components; blockIdx.y = side (0: to rank-1, 1: to rank+1)
__global__ void __launch_bounds__(256)
peer_push_kernel(const double *__restrict__ mine, long long P, const int *__restrict__ src0, const int *__restrict__ dst0, int cnt0, double *nbr0,
                 long long P0, unsigned long long *flag0, const int *__restrict__ src1, const int *__restrict__ dst1, int cnt1, double *nbr1,
                 long long P1, unsigned long long *flag1, unsigned int *counters, unsigned long long seq, const double *__restrict__ scal)
{
    if (scal && scal[7] != 0.0)
        return;
    const int side = blockIdx.y;
    const int *src = side ? src1 : src0, *dst = side ? dst1 : dst0;
    const int cnt = side ? cnt1 : cnt0;
    double *nbr = side ? nbr1 : nbr0;
    const long long PN = side ? P1 : P0;
    if ((side ? flag1 : flag0) == nullptr)
        return;  // no neighbour on this side; with a neighbour but nothing to send the flag is still raised (it waits for it)
    for (int t = blockIdx.x * 256 + threadIdx.x; t < cnt; t += gridDim.x * 256) {
        const long long s = src[t], d = dst[t];
        nbr[d] = mine[s];
        nbr[PN + d] = mine[P + s];
        nbr[2 * PN + d] = mine[2 * P + s];
    }
    // last block of this side raises the neighbour's flag once every block's stores are visible system-wide
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(&counters[side], 1u) + 1u;
        if (done == gridDim.x) {
            counters[side] = 0;  // ready for the next launch (stream-ordered)
            __threadfence_system();
            lpmb_st_release_sys(side ? flag1 : flag0, seq);
        }
    }
}

// index tables: my send rows (brick order) and where they land in the neighbour's vector
int lpmb_peer_halo_setup(lpmb_ctx *c, double *perm_vec, long long P, const int *inv)
{
    if (!lpmb_peer_ready(c))
        return LPMB_OK;
    PeerComm &pc = g_peers[c];
    lpmb_peer_halo_release(c);
    const int own0 = lpmb_own0(c), own1 = lpmb_own1(c);
    const bool lo = c->rank > 0, hi = c->rank < c->world - 1;
    // what I send: inv[] of my boundary rows; what the neighbour tells me: inv[] (its numbering) of its ghost rows
    pc.cnt[0] = lo ? c->narrow_send_lo : 0;
    pc.cnt[1] = hi ? c->narrow_send_hi : 0;
    for (int k = 0; k < 2; k++)
        if (pc.cnt[k] > 0) {
            LPMB_CUDA(cudaMalloc(&pc.src[k], (size_t)pc.cnt[k] * sizeof(int)));
            LPMB_CUDA(cudaMalloc(&pc.dst[k], (size_t)pc.cnt[k] * sizeof(int)));
        }
    if (pc.cnt[0] > 0)
        LPMB_CUDA(cudaMemcpyAsync(pc.src[0], inv + own0, (size_t)pc.cnt[0] * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    if (pc.cnt[1] > 0)
        LPMB_CUDA(cudaMemcpyAsync(pc.src[1], inv + own1 - pc.cnt[1], (size_t)pc.cnt[1] * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    // my lower ghosts [own0-recv_lo, own0) are rank-1's last send_hi owned rows (same order); tell rank-1 where they live here
    LPMB_TRY(lpmb_dist_neighbor_ints(c, inv + own0 - (lo ? c->narrow_recv_lo : 0), lo ? c->narrow_recv_lo : 0, inv + own1, hi ? c->narrow_recv_hi : 0,
                                     pc.dst[0], pc.cnt[0], pc.dst[1], pc.cnt[1]));
    // vector geometry of the neighbours
    long long *d_P;
    LPMB_CUDA(cudaMalloc(&d_P, sizeof(long long) * (c->world + 1)));
    LPMB_CUDA(cudaMemcpyAsync(d_P + c->world, &P, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    LPMB_TRY(lpmb_dist_allgather_bytes(c, d_P + c->world, d_P, sizeof(long long)));
    std::vector<long long> allP(c->world);
    LPMB_CUDA(cudaMemcpyAsync(allP.data(), d_P, sizeof(long long) * c->world, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_P);
    void *opened[LPMB_PEER_MAXW] = {nullptr};
    unsigned want = 0;
    if (lo)
        want |= 1u << (c->rank - 1);
    if (hi)
        want |= 1u << (c->rank + 1);
    bool ok = false;
    LPMB_TRY(exchange_and_open(c, perm_vec, want, opened, &ok));
    bool all_ok = false;
    LPMB_TRY(agree(c, ok, &all_ok));
    if (!all_ok) {
        for (int r = 0; r < c->world; r++)
            if (opened[r] && r != c->rank)
                cudaIpcCloseMemHandle(opened[r]);
        return LPMB_OK;  // halo stays on the NCCL path (all ranks alike)
    }
    if (lo) {
        pc.nbr_vec[0] = (double *)opened[c->rank - 1];
        pc.nbr_P[0] = allP[c->rank - 1];
    }
    if (hi) {
        pc.nbr_vec[1] = (double *)opened[c->rank + 1];
        pc.nbr_P[1] = allP[c->rank + 1];
    }
    pc.my_vec = perm_vec;
    pc.my_P = P;
    pc.halo_ready = true;
    return LPMB_OK;
}

// push my boundary rows of the registered vector into both neighbours; *wait = what my own SpMV must wait for
int lpmb_peer_halo_push(lpmb_ctx *c, const double *scal, PeerWait *wait)
{
    PeerComm &pc = g_peers[c];
    pc.halo_seq++;
    const int maxcnt = pc.cnt[0] > pc.cnt[1] ? pc.cnt[0] : pc.cnt[1];
    int gx = lpmb_blocks(maxcnt, 256);
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    // my flag in the lower neighbour's buffer is its "from rank+1" flag (index 1) and vice versa
    unsigned long long *flag_lo = c->rank > 0 ? reinterpret_cast<unsigned long long *>(pc.peer_buf[c->rank - 1] + OFF_HALO) + 1 : nullptr;
    unsigned long long *flag_hi = c->rank < c->world - 1 ? reinterpret_cast<unsigned long long *>(pc.peer_buf[c->rank + 1] + OFF_HALO) + 0 : nullptr;
    peer_push_kernel<<<dim3(gx, 2), 256, 0, c->stream>>>(pc.my_vec, pc.my_P, pc.src[0], pc.dst[0], pc.cnt[0], pc.nbr_vec[0], pc.nbr_P[0], flag_lo, pc.src[1],
                                                          pc.dst[1], pc.cnt[1], pc.nbr_vec[1], pc.nbr_P[1], flag_hi,
                                                          reinterpret_cast<unsigned int *>(pc.buf + OFF_COUNTER), pc.halo_seq, scal);
    LPMB_LAUNCH_CHECK(c);
    wait->seqs = reinterpret_cast<const unsigned long long *>(pc.buf + OFF_HALO);
    wait->seq = pc.halo_seq;
    wait->n = 2;
    return LPMB_OK;
}

// 0 = single GPU, 1 = NCCL only, 2 = peer-memory scalars (halo through NCCL), 3 = peer-memory scalars and halo push
extern "C" int lpmb_dist_mode(lpmb_ctx *c)
{
    if (!c || c->world <= 1)
        return 0;
    if (!lpmb_peer_ready(c))
        return 1;
    return lpmb_peer_halo_ready(c) ? 3 : 2;
}
