// lpmb_newton.cu -- one pass of the reference's Newton loop body, device resident, plus the
// device-side boundary-condition updates the loop needs when the state never leaves HBM.
//
// Reference: src/lpmc_project.c:426-464
//     switchStateV(0);
//     setDispBC_stiffnessUpdate{2,3}D();        (boundary.c:72-281: zero rows/cols of constrained DoFs in
//                                                K_global, norm_diag on their diagonal, residual := 0)
//     solverCG();                               (solver.c:188-270, then xyz += disp)
//     computeBondForceGeneral(plmode, t);       (+ computeStress + switchStateV(2))
//     updateRR(); norm_residual = dnrm2(residual)
//
// The BC step is realised as a DoF mask inside the solve: constrained DoFs have rhs 0 and x0 = 0, so
// they never enter the Krylov space and the masked operator M K M produces the same iterates as the
// reference's edited matrix (tests/test_solver_gpu.py::test_masked_cg_equals_bc_modified_matrix).  K
// is therefore never touched between assembly and solve and nothing crosses PCIe inside the loop.
//
// setDispBC / setForceBC (boundary.c:12-70) are one add per selected DoF; their device forms below
// keep xyz, dispBC_index and Pex resident (SURVEY section 8(f) "next #2").
#include "lpmb_internal.cuh"

extern "C" int lpmb_newton_iteration(lpmb_ctx *c, int plmode, int load_indicator, double rel, double abs_tol, int maxit, int *cg_iterations,
                                     double *norm_residual)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_TRY(lpmb_switch_state(c, 0));
    LPMB_TRY(lpmb_refresh_mask(c));  // dispBC_index / fix_index may have changed (setDispBC, updateCrack)
    int iters = 0;
    const int rc = lpmb_solve_cg_device(c, rel, abs_tol, maxit, 1, 1, &iters);
    if (rc != LPMB_OK && rc != LPMB_ERR_NOTCONVERGED)
        return rc;
    if (cg_iterations)
        *cg_iterations = iters;
    LPMB_TRY(lpmb_bond_force(c, plmode, load_indicator));
    LPMB_TRY(lpmb_update_rr(c, norm_residual, nullptr));
    return rc;
}

extern "C" int lpmb_field_copy(lpmb_ctx *c, const char *dst, const char *src)
{
    LPMB_REQUIRE(c && dst && src, LPMB_ERR_ARG, "lpmb_field_copy: null argument");
    LPMB_CUDA(cudaSetDevice(c->device));
    Field *d = lpmb_field(c, dst), *s = lpmb_field(c, src);
    LPMB_REQUIRE(d && s, LPMB_ERR_ARG, "lpmb_field_copy: unknown field");
    LPMB_REQUIRE(d->count == s->count && d->elem() == s->elem(), LPMB_ERR_ARG, "lpmb_field_copy %s <- %s: shapes differ", dst, src);
    LPMB_CUDA(cudaMemcpyAsync(d->d, s->d, d->count * d->elem(), cudaMemcpyDeviceToDevice, c->stream));
    return LPMB_OK;
}

// setDispBC (boundary.c:12-45) for one (type, axis, step) entry
__global__ void disp_bc_kernel(int N, int Np, const int *__restrict__ type, int t, int axis, double step, double *__restrict__ xyz,
                               int *__restrict__ bc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || type[i] != t)
        return;
    xyz[(size_t)axis * Np + i] += step;
    bc[(size_t)axis * Np + i] = 0;
}

__global__ void count_type_kernel(int N, const int *__restrict__ type, int t, int *__restrict__ count, int own0, int own1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < N && i >= own0 && i < own1 && type[i] == t;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m)
        atomicAdd(count, __popc(m));
}

// setForceBC (boundary.c:48-70) for one entry: Pex[dim*i+k] += step_k / (number of particles of that type)
__global__ void force_bc_kernel(int N, int Np, int dim, const int *__restrict__ type, int t, double fx, double fy, double fz,
                                double *__restrict__ Pex)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || type[i] != t)
        return;
    Pex[i] += fx;
    Pex[(size_t)Np + i] += fy;
    if (dim == 3)
        Pex[(size_t)2 * Np + i] += fz;
}

extern "C" int lpmb_apply_disp_bc(lpmb_ctx *c, int type, char axis, double step)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    const int ax = axis == 'x' ? 0 : (axis == 'y' ? 1 : (axis == 'z' ? 2 : -1));
    LPMB_REQUIRE(ax >= 0 && ax < c->dim, LPMB_ERR_ARG, "lpmb_apply_disp_bc: axis '%c' with dim %d", axis, c->dim);
    disp_bc_kernel<<<lpmb_blocks(c->N, 256), 256, 0, c->stream>>>(c->N, c->Np, fptr<int>(c, "type"), type, ax, step, fptr<double>(c, "xyz"),
                                                                  fptr<int>(c, "dispBC_index"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}

extern "C" int lpmb_apply_force_bc(lpmb_ctx *c, int type, double step_x, double step_y, double step_z)
{
    LPMB_REQUIRE(c, LPMB_ERR_ARG, "null context");
    LPMB_CUDA(cudaSetDevice(c->device));
    LPMB_TRY(lpmb_cg_alloc(c));
    int *d_count = reinterpret_cast<int *>(c->cg.scal + 14);
    LPMB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
    count_type_kernel<<<lpmb_blocks(c->N, 256), 256, 0, c->stream>>>(c->N, fptr<int>(c, "type"), type, d_count, lpmb_own0(c), lpmb_own1(c));
    LPMB_LAUNCH_CHECK(c);
    int n = 0;
    LPMB_CUDA(cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LPMB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->world > 1) {  // the load is shared by all particles of that type in the whole lattice (boundary.c:53-58)
        double nd = (double)n;
        LPMB_CUDA(cudaMemcpyAsync(c->cg.scal + 9, &nd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LPMB_TRY(lpmb_dist_allreduce_sum(c, c->cg.scal + 9, 1));
        LPMB_CUDA(cudaMemcpyAsync(&nd, c->cg.scal + 9, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        LPMB_CUDA(cudaStreamSynchronize(c->stream));
        n = (int)(nd + 0.5);
    }
    if (n == 0)
        return LPMB_OK;
    // boundary.c:63-66: step / sum_forceBC (double / int)
    force_bc_kernel<<<lpmb_blocks(c->N, 256), 256, 0, c->stream>>>(c->N, c->Np, c->dim, fptr<int>(c, "type"), type, step_x / n, step_y / n,
                                                                   step_z / n, fptr<double>(c, "Pex"));
    LPMB_LAUNCH_CHECK(c);
    return LPMB_OK;
}
