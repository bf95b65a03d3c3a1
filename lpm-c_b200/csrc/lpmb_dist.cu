// lpmb_dist.cu -- particle-slab decomposition across the GPUs of one box: NCCL over NVLink for exactly
// two patterns (north_star): nearest-slab halo exchange and scalar all-reduce.
//
// The reference is a single address space (no MPI/NCCL anywhere); its particle
// order is z-slowest (src/initialization.c:266-284), so contiguous index ranges are z-slabs.  One
// process per GPU holds [ghost_lo | owned | ghost_hi] -- a contiguous sub-lattice in global order, so
// the same topology builder / kernels run unchanged on it.  Ghosts are 4 lattice layers deep: rows in
// the 2 layers next to the owned range have complete 2-hop stars and are computed redundantly (their
// FD-tangent rows are needed for the symmetrisation K_ij = (A_i[j] + A_j[i]^T)/2, their dilatation
// sums and plastic state for the owned bond forces), the outer 2 layers only supply positions.
//   * every CG iteration: the search direction p is exchanged 2 layers deep (conn reach) and the two
//     dot products are all-reduced (8 bytes each);
//   * once per Newton iteration: xyz is exchanged 4 layers deep after xyz += disp;
//   * once per load step: J2_dlambda / J2_triaxiality / damage_nonlocal before / inside the damage update.
// libnccl is dlopen'ed at lpmb_dist_init time (torch's bundled copy when the process already loaded
// it), so single-GPU users of liblpmb200.so carry no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include "lpmb_internal.cuh"

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.lib)
        return LPMB_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
        h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    LPMB_REQUIRE(h, LPMB_ERR_UNSUPPORTED, "cannot load libnccl.so.2: %s", dlerror());
#define LPMB_SYM(field, name)                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                   \
    LPMB_REQUIRE(g_nccl.field, LPMB_ERR_UNSUPPORTED, "libnccl lacks %s", name)
    LPMB_SYM(GetUniqueId, "ncclGetUniqueId");
    LPMB_SYM(CommInitRank, "ncclCommInitRank");
    LPMB_SYM(CommDestroy, "ncclCommDestroy");
    LPMB_SYM(AllReduce, "ncclAllReduce");
    LPMB_SYM(AllGather, "ncclAllGather");
    LPMB_SYM(Send, "ncclSend");
    LPMB_SYM(Recv, "ncclRecv");
    LPMB_SYM(GroupStart, "ncclGroupStart");
    LPMB_SYM(GroupEnd, "ncclGroupEnd");
    LPMB_SYM(GetErrorString, "ncclGetErrorString");
#undef LPMB_SYM
    g_nccl.lib = h;
    return LPMB_OK;
}

#define LPMB_NCCL(call)                                                                                  \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            lpmb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));    \
            return LPMB_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

extern "C" int lpmb_dist_unique_id(void *id128)
{
    LPMB_REQUIRE(id128, LPMB_ERR_ARG, "null id buffer");
    LPMB_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    LPMB_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)));
    return LPMB_OK;
}

extern "C" int lpmb_dist_init(lpmb_ctx *c, const void *id128, int rank, int world)
{
    LPMB_REQUIRE(c && id128 && world >= 1 && rank >= 0 && rank < world, LPMB_ERR_ARG, "lpmb_dist_init: bad argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    LPMB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    c->nccl = comm;
    c->rank = rank;
    c->world = world;
    // NVLink peer-memory fast path for the per-CG-iteration traffic (lpmb_peer.cu); NCCL stays for the rest
    if (world > 1 && !getenv("LPMB_NO_PEER"))
        LPMB_TRY(lpmb_peer_init(c));
    return LPMB_OK;
}

// byte all-gather / neighbour int exchange used by the peer set-up
int lpmb_dist_allgather_bytes(lpmb_ctx *c, const void *d_send, void *d_recv, size_t bytes_per_rank)
{
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    LPMB_NCCL(g_nccl.AllGather(d_send, d_recv, bytes_per_rank, ncclInt8, reinterpret_cast<ncclComm_t>(c->nccl), c->stream));
    return LPMB_OK;
}

// send `to_lo` ints to rank-1 and `to_hi` ints to rank+1, receive `from_lo` / `from_hi` ints from them
int lpmb_dist_neighbor_ints(lpmb_ctx *c, const int *to_lo, int n_to_lo, const int *to_hi, int n_to_hi, int *from_lo, int n_from_lo, int *from_hi,
                            int n_from_hi)
{
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->nccl);
    LPMB_NCCL(g_nccl.GroupStart());
    if (c->rank > 0) {
        if (n_to_lo > 0)
            LPMB_NCCL(g_nccl.Send(to_lo, n_to_lo, ncclInt32, c->rank - 1, comm, c->stream));
        if (n_from_lo > 0)
            LPMB_NCCL(g_nccl.Recv(from_lo, n_from_lo, ncclInt32, c->rank - 1, comm, c->stream));
    }
    if (c->rank < c->world - 1) {
        if (n_to_hi > 0)
            LPMB_NCCL(g_nccl.Send(to_hi, n_to_hi, ncclInt32, c->rank + 1, comm, c->stream));
        if (n_from_hi > 0)
            LPMB_NCCL(g_nccl.Recv(from_hi, n_from_hi, ncclInt32, c->rank + 1, comm, c->stream));
    }
    LPMB_NCCL(g_nccl.GroupEnd());
    return LPMB_OK;
}

void lpmb_dist_release(lpmb_ctx *c)
{
    lpmb_peer_release(c);
    if (c->nccl && g_nccl.CommDestroy)
        g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(c->nccl));
    c->nccl = nullptr;
}

extern "C" int lpmb_dist_set_slab(lpmb_ctx *c, int own0, int own1, int narrow_recv_lo, int narrow_recv_hi, int narrow_send_lo,
                                  int narrow_send_hi, int wide_send_lo, int wide_send_hi)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_REQUIRE(0 <= own0 && own0 < own1 && own1 <= c->N, LPMB_ERR_ARG, "owned range [%d,%d) outside 0..%d", own0, own1, c->N);
    LPMB_REQUIRE(narrow_recv_lo <= own0 && narrow_recv_hi <= c->N - own1, LPMB_ERR_ARG, "narrow halo larger than the ghost region");
    LPMB_REQUIRE(narrow_send_lo <= own1 - own0 && narrow_send_hi <= own1 - own0 && wide_send_lo <= own1 - own0 && wide_send_hi <= own1 - own0,
                 LPMB_ERR_ARG, "send counts exceed the owned range");
    c->own0 = own0;
    c->own1 = own1;
    c->narrow_recv_lo = narrow_recv_lo;
    c->narrow_recv_hi = narrow_recv_hi;
    c->narrow_send_lo = narrow_send_lo;
    c->narrow_send_hi = narrow_send_hi;
    c->wide_send_lo = wide_send_lo;
    c->wide_send_hi = wide_send_hi;
    return LPMB_OK;
}

// Halo exchange of a component-major [comps][Np] fp64 array with rank-1 and rank+1.
//   to rank-1: my first send_lo owned particles  -> its upper ghosts (nearest to its owned range)
//   to rank+1: my last  send_hi owned particles  -> its lower ghosts
int lpmb_dist_exchange(lpmb_ctx *c, double *v, int comps, bool wide)
{
    if (c->world <= 1)
        return LPMB_OK;
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->nccl);
    const int own0 = lpmb_own0(c), own1 = lpmb_own1(c);
    const int recv_lo = wide ? own0 : c->narrow_recv_lo, recv_hi = wide ? c->N - own1 : c->narrow_recv_hi;
    const int send_lo = wide ? c->wide_send_lo : c->narrow_send_lo, send_hi = wide ? c->wide_send_hi : c->narrow_send_hi;
    LPMB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < comps; k++) {
        double *base = v + (size_t)k * c->Np;
        if (c->rank > 0) {
            if (send_lo > 0)
                LPMB_NCCL(g_nccl.Send(base + own0, send_lo, ncclDouble, c->rank - 1, comm, c->stream));
            if (recv_lo > 0)
                LPMB_NCCL(g_nccl.Recv(base + own0 - recv_lo, recv_lo, ncclDouble, c->rank - 1, comm, c->stream));
        }
        if (c->rank < c->world - 1) {
            if (send_hi > 0)
                LPMB_NCCL(g_nccl.Send(base + own1 - send_hi, send_hi, ncclDouble, c->rank + 1, comm, c->stream));
            if (recv_hi > 0)
                LPMB_NCCL(g_nccl.Recv(base + own1, recv_hi, ncclDouble, c->rank + 1, comm, c->stream));
        }
    }
    LPMB_NCCL(g_nccl.GroupEnd());
    return LPMB_OK;
}

int lpmb_dist_allreduce_sum(lpmb_ctx *c, double *d_buf, int count)
{
    if (c->world <= 1)
        return LPMB_OK;
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    LPMB_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(c->nccl), c->stream));
    return LPMB_OK;
}

// contiguous fp64 ranges with rank-1 ("lo") and rank+1 ("hi"): `comps` components `stride` apart, n elements each way
// (multigrid levels of the fast mode: two lattice layers of a level's vector)
int lpmb_dist_neighbor_doubles(lpmb_ctx *c, double *base, long long stride, int comps, long long send_lo_off, long long recv_lo_off,
                               long long send_hi_off, long long recv_hi_off, long long n)
{
    if (c->world <= 1 || n <= 0)
        return LPMB_OK;
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->nccl);
    LPMB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < comps; k++) {
        double *b = base + (size_t)k * stride;
        if (c->rank > 0) {
            LPMB_NCCL(g_nccl.Send(b + send_lo_off, n, ncclDouble, c->rank - 1, comm, c->stream));
            LPMB_NCCL(g_nccl.Recv(b + recv_lo_off, n, ncclDouble, c->rank - 1, comm, c->stream));
        }
        if (c->rank < c->world - 1) {
            LPMB_NCCL(g_nccl.Send(b + send_hi_off, n, ncclDouble, c->rank + 1, comm, c->stream));
            LPMB_NCCL(g_nccl.Recv(b + recv_hi_off, n, ncclDouble, c->rank + 1, comm, c->stream));
        }
    }
    LPMB_NCCL(g_nccl.GroupEnd());
    return LPMB_OK;
}

// every rank contributes the range [offs[rank], offs[rank] + counts[rank]) of each component and ends up with all ranges
// (variable-length all-gather in place; the first replicated multigrid level)
int lpmb_dist_allgatherv_doubles(lpmb_ctx *c, double *base, long long stride, int comps, const long long *offs, const long long *counts)
{
    if (c->world <= 1)
        return LPMB_OK;
    LPMB_REQUIRE(c->nccl, LPMB_ERR_STATE, "lpmb_dist_init not called");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->nccl);
    LPMB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < comps; k++) {
        double *b = base + (size_t)k * stride;
        for (int r = 0; r < c->world; r++) {
            if (r == c->rank)
                continue;
            if (counts[c->rank] > 0)
                LPMB_NCCL(g_nccl.Send(b + offs[c->rank], counts[c->rank], ncclDouble, r, comm, c->stream));
            if (counts[r] > 0)
                LPMB_NCCL(g_nccl.Recv(b + offs[r], counts[r], ncclDouble, r, comm, c->stream));
        }
    }
    LPMB_NCCL(g_nccl.GroupEnd());
    return LPMB_OK;
}

// exchange a named per-particle / DoF fp64 field (harness + damage path)
extern "C" int lpmb_dist_exchange_field(lpmb_ctx *c, const char *name, int wide)
{
    LPMB_REQUIRE(c && name, LPMB_ERR_ARG, "null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    Field *f = lpmb_field(c, name);
    LPMB_REQUIRE(f && f->type == FT_F64 && f->kind != FK_RAW, LPMB_ERR_ARG, "field %s cannot be exchanged", name);
    return lpmb_dist_exchange(c, (double *)f->d, f->comps, wide != 0);
}
