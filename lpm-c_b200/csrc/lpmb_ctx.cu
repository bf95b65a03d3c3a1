// lpmb_ctx.cu -- context lifetime, named-field registry, host<->device re-layout.
//
// The reference keeps every array as a jagged host global (include/lpm.h:55-81, allocated in
// initMatrices, src/initialization.c:1120-1194).  Here each of them is one flat device array in a
// slot-major / component-major layout (DESIGN.md "Data layout in HBM"); lpmb_field_set/get move
// whole arrays and transpose on the device.
#include "lpmb_internal.cuh"

static thread_local char g_err[1024] = "";

void lpmb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *lpmb_last_error(void) { return g_err; }
extern "C" int lpmb_version(void) { return 100; }
extern "C" int lpmb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---- field specs ------------------------------------------------------------------------------
struct FieldSpec {
    const char *name;
    FieldKind kind;
    FieldType type;
    int comps;  // 0 -> nn (bond), -1 -> dim (dof)
    double fill;
};

static const FieldSpec kSpecs[] = {
    // per-bond fp64 (initialization.c:1126-1180)
    {"distance", FK_BOND, FT_F64, 0, 0}, {"distance_initial", FK_BOND, FT_F64, 0, 0},
    {"csx", FK_BOND, FT_F64, 0, 0}, {"csy", FK_BOND, FT_F64, 0, 0}, {"csz", FK_BOND, FT_F64, 0, 0},
    {"csx_initial", FK_BOND, FT_F64, 0, 0}, {"csy_initial", FK_BOND, FT_F64, 0, 0},
    {"csz_initial", FK_BOND, FT_F64, 0, 0},
    {"dL", FK_BOND, FT_F64, 0, 0}, {"dL_ave", FK_BOND, FT_F64, 0, 0}, {"ddL", FK_BOND, FT_F64, 0, 0},
    {"ddLp", FK_BOND, FT_F64, 0, 0}, {"Kn", FK_BOND, FT_F64, 0, 0}, {"Tv", FK_BOND, FT_F64, 0, 0},
    {"F", FK_BOND, FT_F64, 0, 0}, {"F_temp", FK_BOND, FT_F64, 0, 0}, {"bond_stress", FK_BOND, FT_F64, 0, 0},
    {"damage_broken", FK_BOND, FT_F64, 0, 1.0}, {"damage_w", FK_BOND, FT_F64, 0, 1.0},
    {"dLp0", FK_BOND, FT_F64, 0, 0}, {"dLp1", FK_BOND, FT_F64, 0, 0}, {"dLp2", FK_BOND, FT_F64, 0, 0},
    {"damage_D0", FK_BOND, FT_F64, 0, 0}, {"damage_D1", FK_BOND, FT_F64, 0, 0},
    // per-bond integer
    {"neighbors", FK_BOND, FT_I32, 0, -1}, {"nsign", FK_BOND, FT_I8, 0, -1},
    {"mirror", FK_BOND, FT_I8, 0, -1}, {"oppslot", FK_BOND, FT_I8, 0, -1},
    // per-particle fp64
    {"xyz", FK_PART, FT_F64, 3, 0}, {"xyz_initial", FK_PART, FT_F64, 3, 0}, {"xyz_temp", FK_PART, FT_F64, 3, 0},
    {"dL_total", FK_PART, FT_F64, 2, 0}, {"TdL_total", FK_PART, FT_F64, 2, 0},
    {"ddL_total", FK_PART, FT_F64, 2, 0}, {"TddL_total", FK_PART, FT_F64, 2, 0},
    {"stress_tensor", FK_PART, FT_F64, 6, 0}, {"strain_tensor", FK_PART, FT_F64, 6, 0},
    {"J2_beta0", FK_PART, FT_F64, 6, 0}, {"J2_beta1", FK_PART, FT_F64, 6, 0}, {"J2_beta2", FK_PART, FT_F64, 6, 0},
    {"J2_alpha0", FK_PART, FT_F64, 1, 0}, {"J2_alpha1", FK_PART, FT_F64, 1, 0}, {"J2_alpha2", FK_PART, FT_F64, 1, 0},
    {"J2_beta_eq0", FK_PART, FT_F64, 1, 0}, {"J2_beta_eq1", FK_PART, FT_F64, 1, 0},
    {"J2_beta_eq2", FK_PART, FT_F64, 1, 0},
    {"damage_local0", FK_PART, FT_F64, 1, 0}, {"damage_local1", FK_PART, FT_F64, 1, 0},
    {"damage_nonlocal0", FK_PART, FT_F64, 1, 0}, {"damage_nonlocal1", FK_PART, FT_F64, 1, 0},
    {"J2_dlambda", FK_PART, FT_F64, 1, 0}, {"J2_stresseq", FK_PART, FT_F64, 1, 0},
    {"J2_stressm", FK_PART, FT_F64, 1, 0}, {"J2_triaxiality", FK_PART, FT_F64, 1, 0},
    {"sigmay", FK_PART, FT_F64, 1, 0}, {"damage_visual", FK_PART, FT_F64, 1, 0},
    // per-particle integer
    {"type", FK_PART, FT_I32, 1, 0}, {"pl_flag", FK_PART, FT_I32, 1, 0}, {"nb", FK_PART, FT_I32, 1, 0},
    {"nb_initial", FK_PART, FT_I32, 1, 0},
    // memo of the crystal-plasticity law: 1 = this particle's increments of the current pass exist (constitutive.c:946-959)
    {"state_v", FK_PART, FT_I32, 1, 0},
    // DoF vectors
    {"residual", FK_DOF, FT_F64, -1, 0}, {"Pex", FK_DOF, FT_F64, -1, 0}, {"Pex_temp", FK_DOF, FT_F64, -1, 0},
    {"disp", FK_DOF, FT_F64, -1, 0},
    {"dispBC_index", FK_DOF, FT_I32, -1, 1}, {"fix_index", FK_DOF, FT_I32, -1, 1},
    {"Pin", FK_PIN, FT_F64, 3, 0},
    // crystal plasticity (slipSysDefine3D, initialization.c:570,817-826): comps -2 = nslipSys, -3 = nslipSys^2
    {"cp_gy0", FK_PART, FT_F64, -2, 0}, {"cp_gy1", FK_PART, FT_F64, -2, 0}, {"cp_gy2", FK_PART, FT_F64, -2, 0},
    {"cp_A_single0", FK_PART, FT_F64, -2, 0}, {"cp_A_single1", FK_PART, FT_F64, -2, 0}, {"cp_A_single2", FK_PART, FT_F64, -2, 0},
    {"cp_A0", FK_PART, FT_F64, 1, 0}, {"cp_A1", FK_PART, FT_F64, 1, 0}, {"cp_A2", FK_PART, FT_F64, 1, 0},
    {"cp_Cab", FK_PART, FT_F64, -3, 0}, {"cp_RSS", FK_PART, FT_F64, -2, 0}, {"cp_Jact", FK_PART, FT_I32, -2, 0},
    {"cp_dgy", FK_PART, FT_F64, -2, 0}, {"cp_dA", FK_PART, FT_F64, 1, 0}, {"cp_dA_single", FK_PART, FT_F64, -2, 0},
    // snapshots used by harnesses to replay an iteration from the same state
    {"xyz_save", FK_PART, FT_F64, 3, 0}, {"residual_save", FK_DOF, FT_F64, -1, 0},
};

static const FieldSpec *find_spec(const char *name)
{
    for (const FieldSpec &s : kSpecs)
        if (strcmp(s.name, name) == 0)
            return &s;
    return nullptr;
}

template <typename T>
__global__ void fill_kernel(T *p, size_t n, T v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        p[i] = v;
}

int lpmb_field_alloc(lpmb_ctx *c, const char *name, FieldKind kind, FieldType type, int comps)
{
    Field f;
    f.kind = kind;
    f.type = type;
    f.comps = comps;
    f.count = (size_t)comps * c->Np;
    LPMB_CUDA(cudaMalloc(&f.d, f.count * f.elem()));
    LPMB_CUDA(cudaMemsetAsync(f.d, 0, f.count * f.elem(), c->stream));
    c->fields[name] = f;
    return LPMB_OK;
}

// Returns the field, allocating (and default-filling) it on first use if it is a known name.
Field *lpmb_field(lpmb_ctx *c, const char *name)
{
    auto it = c->fields.find(name);
    if (it != c->fields.end())
        return &it->second;
    const FieldSpec *s = find_spec(name);
    if (!s) {
        lpmb_set_error("unknown field '%s'", name);
        return nullptr;
    }
    int comps = s->comps;
    if (s->comps == 0)
        comps = c->nn;
    else if (s->comps == -1)
        comps = c->dim;
    else if (s->comps <= -2) {
        const int S = (int)param(c, "nslipSys", 0.0);
        if (S <= 0) {
            lpmb_set_error("field '%s' needs nslipSys (lpmb_set_schmid_tensor) first", name);
            return nullptr;
        }
        comps = s->comps == -2 ? S : S * S;
    }
    if (lpmb_field_alloc(c, name, s->kind, s->type, comps) != LPMB_OK)
        return nullptr;
    Field *f = &c->fields[name];
    if (s->fill != 0.0) {
        int blocks = 4 * c->sm_count;
        if (f->type == FT_F64)
            fill_kernel<double><<<blocks, 256, 0, c->stream>>>((double *)f->d, f->count, s->fill);
        else if (f->type == FT_I32)
            fill_kernel<int><<<blocks, 256, 0, c->stream>>>((int *)f->d, f->count, (int)s->fill);
        else
            fill_kernel<signed char><<<blocks, 256, 0, c->stream>>>((signed char *)f->d, f->count, (signed char)s->fill);
        c->launches++;
    }
    return f;
}

int lpmb_ensure_staging(lpmb_ctx *c, size_t bytes)
{
    if (bytes <= c->staging_bytes)
        return LPMB_OK;
    if (c->staging) {
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        LPMB_CUDA(cudaFree(c->staging));
        c->staging = nullptr;
        c->staging_bytes = 0;
    }
    LPMB_CUDA(cudaMalloc(&c->staging, bytes));
    c->staging_bytes = bytes;
    return LPMB_OK;
}

int lpmb_ensure_h_staging(lpmb_ctx *c, size_t bytes)
{
    if (bytes <= c->h_staging_bytes)
        return LPMB_OK;
    if (c->h_staging) {
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        LPMB_CUDA(cudaFreeHost(c->h_staging));
        c->h_staging = nullptr;
        c->h_staging_bytes = 0;
    }
    LPMB_CUDA(cudaMallocHost(&c->h_staging, bytes));
    c->h_staging_bytes = bytes;
    return LPMB_OK;
}

// ---- re-layout kernels ------------------------------------------------------------------------
// host-order [N][C] (row-major) <-> device-order [C][Np].  One block per 32 particles; the tile
// goes through shared memory so both the global read and the global write are coalesced.
template <typename TH, typename TD>
__global__ void to_device_layout(const TH *__restrict__ in, TD *__restrict__ out, int N, int Np, int C)
{
    extern __shared__ unsigned char smem_raw[];
    TH *tile = reinterpret_cast<TH *>(smem_raw);
    const int i0 = blockIdx.x * 32;
    const int rows = min(32, N - i0);
    const int total = rows * C;
    for (int e = threadIdx.x; e < total; e += blockDim.x)
        tile[e] = in[(size_t)i0 * C + e];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
        const int c = e >> 5, r = e & 31;
        if (r < rows)
            out[(size_t)c * Np + i0 + r] = (TD)tile[r * C + c];
    }
}

template <typename TH, typename TD>
__global__ void to_host_layout(const TD *__restrict__ in, TH *__restrict__ out, int N, int Np, int C)
{
    extern __shared__ unsigned char smem_raw[];
    TH *tile = reinterpret_cast<TH *>(smem_raw);
    const int i0 = blockIdx.x * 32;
    const int rows = min(32, N - i0);
    for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
        const int c = e >> 5, r = e & 31;
        if (r < rows)
            tile[r * C + c] = (TH)in[(size_t)c * Np + i0 + r];
    }
    __syncthreads();
    const int total = rows * C;
    for (int e = threadIdx.x; e < total; e += blockDim.x)
        out[(size_t)i0 * C + e] = tile[e];
}

static size_t host_elem(const Field &f) { return f.type == FT_F64 ? 8 : 4; }  // int8 fields are int on the host

// element-wise variants for rows too wide for a shared-memory tile (set-up data only)
__global__ void wide_to_device(const double *__restrict__ in, double *__restrict__ out, int N, int Np, int C)
{
    const size_t total = (size_t)N * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t comp = e / N, i = e % N;  // consecutive threads -> consecutive particles of one component
        out[comp * Np + i] = in[i * C + comp];
    }
}
__global__ void wide_to_host(const double *__restrict__ in, double *__restrict__ out, int N, int Np, int C)
{
    const size_t total = (size_t)N * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t comp = e / N, i = e % N;
        out[i * C + comp] = in[comp * Np + i];
    }
}

// host [N][comps] doubles -> device [comps][Np] (not a registered field: CG vectors, masks, ...)
int lpmb_upload_soa_f64(lpmb_ctx *c, const double *host, double *d_dst, int comps)
{
    const size_t bytes = (size_t)c->N * comps * 8;
    LPMB_TRY(lpmb_ensure_staging(c, bytes));
    LPMB_CUDA(cudaMemcpyAsync(c->staging, host, bytes, cudaMemcpyHostToDevice, c->stream));
    to_device_layout<double, double><<<c->Np / 32, 256, (size_t)32 * comps * 8, c->stream>>>((const double *)c->staging, d_dst, c->N, c->Np, comps);
    LPMB_LAUNCH_CHECK(c);
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return LPMB_OK;
}

int lpmb_download_soa_f64(lpmb_ctx *c, const double *d_src, double *host, int comps)
{
    const size_t bytes = (size_t)c->N * comps * 8;
    LPMB_TRY(lpmb_ensure_staging(c, bytes));
    to_host_layout<double, double><<<c->Np / 32, 256, (size_t)32 * comps * 8, c->stream>>>(d_src, (double *)c->staging, c->N, c->Np, comps);
    LPMB_LAUNCH_CHECK(c);
    LPMB_CUDA(cudaMemcpyAsync(host, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return LPMB_OK;
}

extern "C" int lpmb_field_set(lpmb_ctx *c, const char *name, const void *host, size_t count)
{
    LPMB_REQUIRE(c && name && host, LPMB_ERR_ARG, "lpmb_field_set: null argument");
    Field *f = lpmb_field(c, name);
    if (!f)
        return LPMB_ERR_ARG;
    if (f->kind == FK_RAW) {
        LPMB_REQUIRE(count == f->count, LPMB_ERR_ARG, "field %s: expected %zu elements, got %zu", name, f->count, count);
        LPMB_CUDA(cudaMemcpyAsync(f->d, host, count * f->elem(), cudaMemcpyHostToDevice, c->stream));
        return LPMB_OK;
    }
    const size_t expect = (size_t)c->N * f->comps;
    LPMB_REQUIRE(count == expect, LPMB_ERR_ARG, "field %s: expected %zu elements, got %zu", name, expect, count);
    const size_t bytes = count * host_elem(*f);
    LPMB_TRY(lpmb_ensure_staging(c, bytes));
    LPMB_CUDA(cudaMemcpyAsync(c->staging, host, bytes, cudaMemcpyHostToDevice, c->stream));
    const int blocks = c->Np / 32;
    const size_t smem = (size_t)32 * f->comps * host_elem(*f);
    if (smem > 40 * 1024) {  // very wide rows (cp_Cab: nslipSys^2 components): plain element-wise re-layout
        LPMB_REQUIRE(f->type == FT_F64, LPMB_ERR_UNSUPPORTED, "wide integer field %s", name);
        wide_to_device<<<4 * c->sm_count, 256, 0, c->stream>>>((const double *)c->staging, (double *)f->d, c->N, c->Np, f->comps);
    } else if (f->type == FT_F64)
        to_device_layout<double, double><<<blocks, 256, smem, c->stream>>>((const double *)c->staging, (double *)f->d, c->N, c->Np, f->comps);
    else if (f->type == FT_I32)
        to_device_layout<int, int><<<blocks, 256, smem, c->stream>>>((const int *)c->staging, (int *)f->d, c->N, c->Np, f->comps);
    else
        to_device_layout<int, signed char><<<blocks, 256, smem, c->stream>>>((const int *)c->staging, (signed char *)f->d, c->N, c->Np, f->comps);
    LPMB_LAUNCH_CHECK(c);
    // the staging buffer is reused by the next call: keep ordering simple
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return LPMB_OK;
}

extern "C" int lpmb_field_get(lpmb_ctx *c, const char *name, void *host, size_t count)
{
    LPMB_REQUIRE(c && name && host, LPMB_ERR_ARG, "lpmb_field_get: null argument");
    Field *f = lpmb_field(c, name);
    if (!f)
        return LPMB_ERR_ARG;
    if (f->kind == FK_RAW) {
        LPMB_REQUIRE(count == f->count, LPMB_ERR_ARG, "field %s: expected %zu elements, got %zu", name, f->count, count);
        LPMB_CUDA(cudaMemcpyAsync(host, f->d, count * f->elem(), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        return LPMB_OK;
    }
    const size_t expect = (size_t)c->N * f->comps;
    LPMB_REQUIRE(count == expect, LPMB_ERR_ARG, "field %s: expected %zu elements, got %zu", name, expect, count);
    const size_t bytes = count * host_elem(*f);
    LPMB_TRY(lpmb_ensure_staging(c, bytes));
    const int blocks = c->Np / 32;
    const size_t smem = (size_t)32 * f->comps * host_elem(*f);
    if (smem > 40 * 1024) {
        LPMB_REQUIRE(f->type == FT_F64, LPMB_ERR_UNSUPPORTED, "wide integer field %s", name);
        wide_to_host<<<4 * c->sm_count, 256, 0, c->stream>>>((const double *)f->d, (double *)c->staging, c->N, c->Np, f->comps);
    } else if (f->type == FT_F64)
        to_host_layout<double, double><<<blocks, 256, smem, c->stream>>>((const double *)f->d, (double *)c->staging, c->N, c->Np, f->comps);
    else if (f->type == FT_I32)
        to_host_layout<int, int><<<blocks, 256, smem, c->stream>>>((const int *)f->d, (int *)c->staging, c->N, c->Np, f->comps);
    else
        to_host_layout<int, signed char><<<blocks, 256, smem, c->stream>>>((const signed char *)f->d, (int *)c->staging, c->N, c->Np, f->comps);
    LPMB_LAUNCH_CHECK(c);
    LPMB_CUDA(cudaMemcpyAsync(host, c->staging, bytes, cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return LPMB_OK;
}

// Several fields with ONE device->host copy and ONE synchronisation: the host-layout images are laid out back to back
// in the context's pinned staging buffer; staged[k] points at the image of names[k] (valid until the next call that
// uses the staging buffers).  The caller copies / scatters from there (the drop-in layer scatters into the reference's
// jagged arrays directly, with no intermediate flat copy).
extern "C" int lpmb_fields_get_staged(lpmb_ctx *c, int n, const char *const *names, const void **staged, size_t *counts)
{
    LPMB_REQUIRE(c && names && staged && n > 0 && n <= 64, LPMB_ERR_ARG, "lpmb_fields_get_staged: bad argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    size_t off[65];
    Field *fs[64];
    off[0] = 0;
    for (int k = 0; k < n; k++) {
        fs[k] = lpmb_field(c, names[k]);
        if (!fs[k])
            return LPMB_ERR_ARG;
        const size_t cnt = fs[k]->kind == FK_RAW ? fs[k]->count : (size_t)c->N * fs[k]->comps;
        const size_t bytes = cnt * (fs[k]->kind == FK_RAW ? fs[k]->elem() : host_elem(*fs[k]));
        if (counts)
            counts[k] = cnt;
        off[k + 1] = off[k] + ((bytes + 255) & ~(size_t)255);
    }
    LPMB_TRY(lpmb_ensure_staging(c, off[n]));
    LPMB_TRY(lpmb_ensure_h_staging(c, off[n]));
    for (int k = 0; k < n; k++) {
        Field *f = fs[k];
        char *dst = (char *)c->staging + off[k];
        if (f->kind == FK_RAW) {
            LPMB_CUDA(cudaMemcpyAsync(dst, f->d, f->count * f->elem(), cudaMemcpyDeviceToDevice, c->stream));
            continue;
        }
        const int blocks = c->Np / 32;
        const size_t smem = (size_t)32 * f->comps * host_elem(*f);
        if (smem > 40 * 1024) {
            LPMB_REQUIRE(f->type == FT_F64, LPMB_ERR_UNSUPPORTED, "wide integer field %s", names[k]);
            wide_to_host<<<4 * c->sm_count, 256, 0, c->stream>>>((const double *)f->d, (double *)dst, c->N, c->Np, f->comps);
        } else if (f->type == FT_F64)
            to_host_layout<double, double><<<blocks, 256, smem, c->stream>>>((const double *)f->d, (double *)dst, c->N, c->Np, f->comps);
        else if (f->type == FT_I32)
            to_host_layout<int, int><<<blocks, 256, smem, c->stream>>>((const int *)f->d, (int *)dst, c->N, c->Np, f->comps);
        else
            to_host_layout<int, signed char><<<blocks, 256, smem, c->stream>>>((const signed char *)f->d, (int *)dst, c->N, c->Np, f->comps);
        LPMB_LAUNCH_CHECK(c);
    }
    LPMB_CUDA(cudaMemcpyAsync(c->h_staging, c->staging, off[n], cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n; k++)
        staged[k] = (const char *)c->h_staging + off[k];
    return LPMB_OK;
}

extern "C" int lpmb_field_device(lpmb_ctx *c, const char *name, void **dptr, size_t *count)
{
    LPMB_REQUIRE(c && name, LPMB_ERR_ARG, "lpmb_field_device: null argument");
    Field *f = lpmb_field(c, name);
    if (!f)
        return LPMB_ERR_ARG;
    if (dptr)
        *dptr = f->d;
    if (count)
        *count = f->count;
    return LPMB_OK;
}

// ---- context ----------------------------------------------------------------------------------
extern "C" int lpmb_create(lpmb_ctx **out, int device, int nparticle, int dim, int lattice, int nneighbors, int nconn_max)
{
    LPMB_REQUIRE(out, LPMB_ERR_ARG, "lpmb_create: null out pointer");
    *out = nullptr;
    LPMB_REQUIRE(nparticle > 0 && (dim == 2 || dim == 3) && nneighbors > 0 && nneighbors <= 32 && nconn_max > 0 && nconn_max <= 128,
                 LPMB_ERR_ARG, "lpmb_create: bad sizes (N=%d dim=%d nn=%d nconn=%d)", nparticle, dim, nneighbors, nconn_max);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        lpmb_set_error("lpmb_create: no CUDA device available (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return LPMB_ERR_CUDA;
    }
    LPMB_REQUIRE(device >= 0 && device < ndev, LPMB_ERR_ARG, "lpmb_create: device %d out of range (%d devices)", device, ndev);
    LPMB_CUDA(cudaSetDevice(device));
    lpmb_ctx *c = new lpmb_ctx();
    c->device = device;
    c->N = nparticle;
    c->Np = (nparticle + LPMB_SLICE - 1) / LPMB_SLICE * LPMB_SLICE;
    c->dim = dim;
    c->lattice = lattice;
    c->nn = nneighbors;
    c->nconn = nconn_max;
    c->K.D = dim;
    cudaDeviceProp prop;
    LPMB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    LPMB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    return LPMB_OK;
}

extern "C" void lpmb_destroy(lpmb_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    lpmb_grid_release(c);
    lpmb_dist_release(c);
    lpmb_brick_release(c);
    for (auto &e : c->prof_events)
        cudaEventDestroy(e);
    for (auto &kv : c->fields)
        cudaFree(kv.second.d);
    cudaFree(c->K.sptr);
    cudaFree(c->K.col);
    cudaFree(c->K.val);
    cudaFree(c->K.nbc);
    cudaFree(c->K.k0);
    cudaFree(c->K.kp);
    cudaFree(c->cg.r);
    cudaFree(c->cg.p);
    cudaFree(c->cg.ap);
    cudaFree(c->cg.x);
    lpmb_mg_release(c);
    cudaFree(c->cg.z);
    cudaFree(c->K.rptr);
    cudaFree(c->K.rcol);
    cudaFree(c->K.rval);
    cudaFree(c->cg.partials);
    cudaFree(c->cg.scal);
    cudaFree(c->cg.counters);
    if (c->cg.graph)
        cudaGraphExecDestroy(c->cg.graph);
    if (c->cg.h_scal)
        cudaFreeHost(c->cg.h_scal);
    cudaFree(c->mask);
    cudaFree(c->fd_tab);
    cudaFree(c->staging);
    if (c->h_staging)
        cudaFreeHost(c->h_staging);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int lpmb_synchronize(lpmb_ctx *c)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    return LPMB_OK;
}

extern "C" long long lpmb_launch_count(lpmb_ctx *c) { return c ? c->launches : 0; }
extern "C" void *lpmb_stream(lpmb_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int lpmb_set_param(lpmb_ctx *c, const char *name, double value)
{
    LPMB_REQUIRE(c && name, LPMB_ERR_ARG, "lpmb_set_param: null argument");
    c->params[name] = value;
    return LPMB_OK;
}

extern "C" int lpmb_get_param(lpmb_ctx *c, const char *name, double *value)
{
    LPMB_REQUIRE(c && name && value, LPMB_ERR_ARG, "lpmb_get_param: null argument");
    auto it = c->params.find(name);
    LPMB_REQUIRE(it != c->params.end(), LPMB_ERR_ARG, "parameter '%s' not set", name);
    *value = it->second;
    return LPMB_OK;
}
