// lpmb_io.cu -- compact binary snapshots of a device context (checkpoint / resume).
//
// The reference has no restart files: its drivers dump O(N) TEXT every load step (data_handler.c:42-84, ~2 GB per
// step at 10M particles) and keep roll-back state only in memory.  A run on
// a >= 10M-particle lattice needs a compact record instead: one file = the context's scalar parameters + every
// registered field in its device layout ([comps][Np], exactly the bytes in HBM: no re-layout, chunks go through the
// pinned staging buffer).  Loading restores a context of the same shape bit for bit; what is derived from fields
// (block pattern of K, DoF mask) is rebuilt, the tangent VALUES are not stored (46.8 GB at 216^3 against 0.83 s for
// one assembly, which every load step starts with anyway).  Host code only -- no kernels of its own.
//
// File layout (little endian):
//   char[8]  "LPMBSNP1"
//   int32    N, Np, dim, lattice, nn, nconn, nparams, nfields
//   nparams  x { uint16 len; char name[len]; double value }     (slab runs add "__world", "__rank", "__own0", "__own1")
//   nfields  x { uint16 len; char name[len]; int32 kind, type, comps; uint64 count; byte data[count * elem] }
#include <cstdio>

#include "lpmb_internal.cuh"

int lpmb_rebuild_connectivity(lpmb_ctx *c);  // lpmb_topology.cu

static const char SNAP_MAGIC[8] = {'L', 'P', 'M', 'B', 'S', 'N', 'P', '1'};
static const size_t SNAP_CHUNK = (size_t)64 << 20;

// scratch that is not state: twins of the per-particle laws, plmode-5 work arrays, previous-value copies
static bool snap_skip(const std::string &name)
{
    return name.rfind("pp.", 0) == 0 || name.rfind("iso_", 0) == 0 || name == "dL_prev" || name == "dL_total_prev" || name == "TdL_total_prev";
}

struct FileCloser {
    FILE *f;
    ~FileCloser()
    {
        if (f)
            fclose(f);
    }
};

#define SNAP_IO(cond, what)                                                    \
    do {                                                                       \
        if (!(cond)) {                                                         \
            lpmb_set_error("snapshot %s: %s failed", path, what);              \
            return LPMB_ERR_ARG;                                               \
        }                                                                      \
    } while (0)

extern "C" int lpmb_snapshot_save(lpmb_ctx *c, const char *path)
{
    LPMB_REQUIRE(c && path, LPMB_ERR_ARG, "lpmb_snapshot_save: null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    FileCloser fc{fopen(path, "wb")};
    SNAP_IO(fc.f, "open for writing");
    int nfields = 0;
    for (auto &kv : c->fields)
        if (!snap_skip(kv.first))
            nfields++;
    // slab runs: the file is this rank's slab; its place in the decomposition travels as four pseudo-parameters
    std::map<std::string, double> params = c->params;
    if (c->world > 1) {
        params["__world"] = c->world;
        params["__rank"] = c->rank;
        params["__own0"] = lpmb_own0(c);
        params["__own1"] = lpmb_own1(c);
    }
    const int head[8] = {c->N, c->Np, c->dim, c->lattice, c->nn, c->nconn, (int)params.size(), nfields};
    SNAP_IO(fwrite(SNAP_MAGIC, 1, 8, fc.f) == 8 && fwrite(head, sizeof(int), 8, fc.f) == 8, "write header");
    for (auto &kv : params) {
        const unsigned short len = (unsigned short)kv.first.size();
        SNAP_IO(fwrite(&len, 2, 1, fc.f) == 1 && fwrite(kv.first.data(), 1, len, fc.f) == len && fwrite(&kv.second, 8, 1, fc.f) == 1, "write parameter");
    }
    LPMB_TRY(lpmb_ensure_h_staging(c, SNAP_CHUNK));
    for (auto &kv : c->fields) {
        if (snap_skip(kv.first))
            continue;
        const Field &f = kv.second;
        const unsigned short len = (unsigned short)kv.first.size();
        const int meta[3] = {(int)f.kind, (int)f.type, f.comps};
        const unsigned long long count = f.count;
        SNAP_IO(fwrite(&len, 2, 1, fc.f) == 1 && fwrite(kv.first.data(), 1, len, fc.f) == len && fwrite(meta, sizeof(int), 3, fc.f) == 3 &&
                    fwrite(&count, 8, 1, fc.f) == 1,
                "write field header");
        const size_t bytes = f.count * f.elem();
        for (size_t off = 0; off < bytes; off += SNAP_CHUNK) {
            const size_t n = bytes - off < SNAP_CHUNK ? bytes - off : SNAP_CHUNK;
            LPMB_CUDA(cudaMemcpyAsync(c->h_staging, (const char *)f.d + off, n, cudaMemcpyDeviceToHost, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            SNAP_IO(fwrite(c->h_staging, 1, n, fc.f) == n, "write field data");
        }
    }
    SNAP_IO(fflush(fc.f) == 0, "flush");
    return LPMB_OK;
}

extern "C" int lpmb_snapshot_load(lpmb_ctx *c, const char *path)
{
    LPMB_REQUIRE(c && path, LPMB_ERR_ARG, "lpmb_snapshot_load: null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    // slab runs: one file per rank, loaded AFTER lpmb_dist_init + lpmb_dist_set_slab into a context with the same
    // decomposition (checked below against the "__world/__rank/__own0/__own1" entries the save wrote)
    FileCloser fc{fopen(path, "rb")};
    SNAP_IO(fc.f, "open for reading");
    char magic[8];
    int head[8];
    SNAP_IO(fread(magic, 1, 8, fc.f) == 8 && memcmp(magic, SNAP_MAGIC, 8) == 0, "magic check");
    SNAP_IO(fread(head, sizeof(int), 8, fc.f) == 8, "read header");
    LPMB_REQUIRE(head[0] == c->N && head[1] == c->Np && head[2] == c->dim && head[3] == c->lattice && head[4] == c->nn && head[5] == c->nconn,
                 LPMB_ERR_ARG, "snapshot %s is for N=%d dim=%d lattice=%d nn=%d nconn=%d, the context has N=%d dim=%d lattice=%d nn=%d nconn=%d", path,
                 head[0], head[2], head[3], head[4], head[5], c->N, c->dim, c->lattice, c->nn, c->nconn);
    char name[256];
    int s_world = 1, s_rank = 0, s_own0 = 0, s_own1 = c->N;
    for (int k = 0; k < head[6]; k++) {
        unsigned short len = 0;
        double v = 0;
        SNAP_IO(fread(&len, 2, 1, fc.f) == 1 && len < sizeof(name) && fread(name, 1, len, fc.f) == len && fread(&v, 8, 1, fc.f) == 1, "read parameter");
        name[len] = 0;
        if (name[0] == '_' && name[1] == '_') {
            const std::string key(name);
            if (key == "__world")
                s_world = (int)v;
            else if (key == "__rank")
                s_rank = (int)v;
            else if (key == "__own0")
                s_own0 = (int)v;
            else if (key == "__own1")
                s_own1 = (int)v;
            continue;
        }
        c->params[name] = v;
    }
    LPMB_REQUIRE(s_world == c->world && s_rank == c->rank && (c->world == 1 || (s_own0 == lpmb_own0(c) && s_own1 == lpmb_own1(c))), LPMB_ERR_ARG,
                 "snapshot %s was written by rank %d of %d (owned particles [%d, %d)); this context is rank %d of %d (owned [%d, %d)) -- slab "
                 "snapshots are per rank and need the same decomposition (lpmb_dist_init + lpmb_dist_set_slab before the load)",
                 path, s_rank, s_world, s_own0, s_own1, c->rank, c->world, lpmb_own0(c), lpmb_own1(c));
    LPMB_TRY(lpmb_ensure_h_staging(c, SNAP_CHUNK));
    for (int k = 0; k < head[7]; k++) {
        unsigned short len = 0;
        int meta[3];
        unsigned long long count = 0;
        SNAP_IO(fread(&len, 2, 1, fc.f) == 1 && len < sizeof(name) && fread(name, 1, len, fc.f) == len && fread(meta, sizeof(int), 3, fc.f) == 3 &&
                    fread(&count, 8, 1, fc.f) == 1,
                "read field header");
        name[len] = 0;
        auto it = c->fields.find(name);
        if (it != c->fields.end() && (it->second.count != count || (int)it->second.type != meta[1] || it->second.comps != meta[2])) {
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(it->second.d);
            c->fields.erase(it);
            it = c->fields.end();
        }
        if (it == c->fields.end()) {
            LPMB_REQUIRE(meta[0] >= FK_BOND && meta[0] <= FK_RAW && meta[1] >= FT_F64 && meta[1] <= FT_I8 && meta[2] > 0 && count > 0 &&
                             (meta[0] == FK_RAW ? count < ((unsigned long long)1 << 40) : count == (unsigned long long)meta[2] * c->Np),
                         LPMB_ERR_ARG, "snapshot %s: field %s has an unexpected shape", path, name);
            if (meta[0] == FK_RAW) {  // small tables (Ce, KnTve, schmid_tensor): element count is not tied to Np
                Field nf;
                nf.kind = FK_RAW;
                nf.type = (FieldType)meta[1];
                nf.comps = meta[2];
                nf.count = (size_t)count;
                LPMB_CUDA(cudaMalloc(&nf.d, nf.count * nf.elem()));
                c->fields[name] = nf;
            } else {
                LPMB_TRY(lpmb_field_alloc(c, name, (FieldKind)meta[0], (FieldType)meta[1], meta[2]));
            }
            it = c->fields.find(name);
        }
        Field &f = it->second;
        const size_t bytes = f.count * f.elem();
        for (size_t off = 0; off < bytes; off += SNAP_CHUNK) {
            const size_t n = bytes - off < SNAP_CHUNK ? bytes - off : SNAP_CHUNK;
            SNAP_IO(fread(c->h_staging, 1, n, fc.f) == n, "read field data");
            LPMB_CUDA(cudaMemcpyAsync((char *)f.d + off, c->h_staging, n, cudaMemcpyHostToDevice, c->stream));
            LPMB_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    // derived structures: block pattern of K from the neighbour lists (values are re-assembled by the caller),
    // DoF mask from the BC index fields
    c->K.values_ready = false;
    const bool had_bricks = lpmb_brick_active(c);
    lpmb_brick_touch(c);
    if (c->fields.count("neighbors") && c->fields.count("nsign")) {
        LPMB_TRY(lpmb_rebuild_connectivity(c));   // releases the brick mirror of the previous pattern
        if (had_bricks)                           // ... which the resumed run would silently lose (CG on the full-format kernel)
            LPMB_TRY(lpmb_matrix_enable_bricks(c, 1));
    }
    if (c->fields.count("dispBC_index") && c->fields.count("fix_index"))
        LPMB_TRY(lpmb_refresh_mask(c));
    return LPMB_OK;
}
