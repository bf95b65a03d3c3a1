/*
 * lpmb200.h -- C ABI of the B200-native LPM hot path (liblpmb200.so).
 *
 * Plain C: opaque context, plain pointers and sizes, int status codes (0 = ok; the text of
 * the last error is returned by lpmb_last_error()).  There is NO CPU fallback: every compute
 * entry point fails (non-zero) when no CUDA device is usable.
 *
 * What each group replaces in the reference (ymlasu/LPM-C, paths relative to its root):
 *
 *   topology / CSR container   src/neighbor.c:9-141 (searchNormalNeighbor, searchAFEMNeighbor:
 *                              neighbors, nsign, conn, nb_conn, K_pointer, IK/JK sizing)
 *   matrix import/export       the 3-array 1-based symmetric-upper CSR K_global/IK/JK that
 *                              src/stiffness.c:441-515 writes and src/solver.c:206 consumes
 *   lpmb_solve_cg              solverCG(), src/solver.c:188-270 (MKL RCI dcg + mkl_sparse_d_mv)
 *   lpmb_fd_stiffness          calcStiffness{2,3}DFiniteDifference(6), src/stiffness.c:271-516
 *   lpmb_calc_kntv             calcKnTv(), src/stiffness.c:11-268
 *   lpmb_bond_force            computeBondForceGeneral(plmode,t), src/constitutive.c:88-146
 *                              (+ computeStress, src/lpm_basic.c:53-125, + switchStateV(2))
 *   lpmb_compute_dl            computedL(), src/lpm_basic.c:252-291
 *   lpmb_switch_state          switchStateV(flag), src/constitutive.c:10-85
 *   lpmb_update_rr             updateRR(), src/stiffness.c:519-534
 *   lpmb_update_damage         updateDamageGeneral(), src/constitutive.c:149-164 ->
 *                              updateDuctileDamagePwiseNonlocal (:1757-1862), updateBrittleDamage (:1437-1526),
 *                              updateDuctileDamageBwiseLocal (:1607-1695), updateDuctileDamagePwiseLocal (:1529-1579),
 *                              updateDuctileDamageBwiseNonlocal (:1698-1753)
 *   lpmb_update_crack          updateCrack(), src/constitutive.c:1399-1434
 *   lpmb_set_dof_mask          the effect of setDispBC_stiffnessUpdate{2,3}D, src/boundary.c:72-281,
 *                              as a DoF mask applied inside the solve (K is never edited)
 *   lpmb_apply_disp_bc / _force_bc   setDispBC / setForceBC, src/boundary.c:12-70, on the resident arrays
 *   lpmb_bond_force_particle   computeBondForceElastic / IncrementalUpdating / J2mixedLinear3D / J2energyReturnMap /
 *                              J2nonlinearIso / CPMiehe(ii), src/constitutive.c:167-283, 286-463, 466-686, 689-863, 866-1396
 *   lpmb_compute_strain        computeStrain(), src/lpm_basic.c:127-249
 *   lpmb_compute_cab, lpmb_set_schmid_tensor   computeCab(), src/constitutive.c:1864-1917; plmode 1 of
 *                              lpmb_bond_force = computeBondForceCPMiehe, :866-1396
 *   lpmb_newton_iteration      one pass of the driver's Newton loop, src/lpmc_project.c:426-464
 *   lpmb_snapshot_save / _load no counterpart (the reference has text dumps only, data_handler.c:42-84)
 *
 * The reference-named drop-in entry points (void solverCG(void) ... on the reference's process
 * globals) live in liblpmc_dropin (lpm-c_b200/csrc/dropin.c) and are thin wrappers over this ABI;
 * INTEGRATION.md shows how the reference links against them.
 *
 * Host array layouts are the reference's logical layouts, flattened row-major:
 *   per-bond      [nparticle][nneighbors]            (reference: T **a, a[i][j])
 *   per-particle  [nparticle][c]                      (xyz: c=3, stress_tensor: c=6, ...)
 *   DoF vectors   [nparticle*dim] interleaved         (residual, Pex, disp), Pin: [nparticle*3]
 * Device layouts are slot-major / component-major -- conversion happens on the
 * device inside lpmb_field_set / lpmb_field_get.
 */
#ifndef LPMB200_H
#define LPMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpmb_ctx lpmb_ctx;

#define LPMB_OK 0
#define LPMB_ERR_CUDA 1
#define LPMB_ERR_ARG 2
#define LPMB_ERR_STATE 3
#define LPMB_ERR_NOTCONVERGED 4
#define LPMB_ERR_UNSUPPORTED 5

/* model-level constants mirrored from include/lpm.h:34-51 */
#define LPMB_EPS 1e-6      /* EPS: FD perturbation coefficient / "is zero" threshold */
#define LPMB_TOLITER 1e-4  /* TOLITER */

/* lattice ids, src/lpmc_project.c:70-75 */
enum { LPMB_LATTICE_SQUARE = 0, LPMB_LATTICE_HEX = 1, LPMB_LATTICE_SC = 2, LPMB_LATTICE_FCC = 3, LPMB_LATTICE_BCC = 4 };

const char *lpmb_last_error(void);
int lpmb_version(void);
/* number of usable CUDA devices (0 on a CPU-only host; never an error) */
int lpmb_device_count(void);

/* ---- context ------------------------------------------------------------------------------- */
/* nconn_max = nneighbors_AFEM + 1 (row length of the reference's conn[][]). */
int lpmb_create(lpmb_ctx **out, int device, int nparticle, int dim, int lattice, int nneighbors, int nconn_max);
void lpmb_destroy(lpmb_ctx *ctx);
int lpmb_synchronize(lpmb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long lpmb_launch_count(lpmb_ctx *ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on */
void *lpmb_stream(lpmb_ctx *ctx);

/* scalar parameters by reference name: radius, particle_volume, J2_H, J2_xi, damage_L,
 * damage_threshold, damagec_A, damageb_A, critical_bstrain, dtime, ... ; ints: nbreak, plmode.
 * Implementation switches: cg_precond, fd_variant, fd_tab_mb, cp_warp, cg_graph, j2_fused,
 * spmv_rows, spmv_rows_max, brick_trim, brick_lazy_wait, peer_comm. */
int lpmb_set_param(lpmb_ctx *ctx, const char *name, double value);
int lpmb_get_param(lpmb_ctx *ctx, const char *name, double *value);

/* ---- fields (named arrays; names are the reference's global names) ------------------------- */
/* host -> device / device -> host, host layout as documented above.  `count` = number of
 * elements in the host array (checked).  3-slot state arrays are addressed per slot:
 * "dLp0","dLp1","dLp2","J2_alpha0".., "J2_beta0"..(6 comps), "damage_D0","damage_D1",
 * "damage_nonlocal0/1", "damage_local0/1". */
int lpmb_field_set(lpmb_ctx *ctx, const char *name, const void *host, size_t count);
int lpmb_field_get(lpmb_ctx *ctx, const char *name, void *host, size_t count);
/* n <= 64 fields with ONE device->host copy and ONE synchronisation: staged[k] points at the host-layout image of
 * names[k] inside the context's pinned staging buffer (valid until the next call that stages data), counts[k] (optional)
 * receives its element count.  What liblpmc_dropin.so uses to refresh the reference's host arrays after a call: it
 * scatters straight from the pinned images into the jagged arrays. */
int lpmb_fields_get_staged(lpmb_ctx *ctx, int n, const char *const *names, const void **staged, size_t *counts);
/* raw device pointer + element count of a field (device layout; for zero-copy harnesses) */
int lpmb_field_device(lpmb_ctx *ctx, const char *name, void **dptr, size_t *count);

/* ---- topology ------------------------------------------------------------------------------ */
/* Uploads neighbors/nsign ([N][nn], -1 padded, ascending j as neighbor.c:16-41 produces) and
 * derives nb_initial, mirror slots and opposite-bond slots on the device. */
int lpmb_set_neighbors(lpmb_ctx *ctx, const int *neighbors, const int *nsign);
/* Uploads conn ([N][nconn_max], sorted ascending, -1 padded) and builds the block pattern,
 * K_pointer (64-bit offsets) and the SELL-32 slices. */
int lpmb_set_connectivity(lpmb_ctx *ctx, const int *conn);
/* O(N) device builder of the same lists from xyz (cell grid): bit-identical to
 * searchNormalNeighbor + searchAFEMNeighbor for lattices without duplicates.  Also fills
 * distance_initial, cs{x,y,z}_initial.  cutoff1/cutoff2 = neighbor1_cutoff/neighbor2_cutoff. */
int lpmb_build_topology(lpmb_ctx *ctx, double cutoff1, double cutoff2);
/* sizes of the reference CSR for this connectivity */
int lpmb_csr_sizes(lpmb_ctx *ctx, long long *nnz_upper, long long *nblocks);
/* K_pointer as the reference lays it out: [N+1][2] ints (fails if nnz_upper > INT_MAX) */
int lpmb_get_k_pointer(lpmb_ctx *ctx, int *k_pointer);

/* ---- stiffness matrix ---------------------------------------------------------------------- */
/* import the reference's symmetric-upper 1-based CSR (host arrays) into the device matrix */
int lpmb_matrix_from_upper_csr(lpmb_ctx *ctx, const double *K_global, long long nnz);
/* export the device matrix as the reference's K_global / IK / JK (any pointer may be NULL) */
int lpmb_matrix_to_upper_csr(lpmb_ctx *ctx, double *K_global, int *IK, int *JK);
/* FD elastic tangent (forward difference, h = EPS*radius), symmetrised as stiffness.c:441-481.
 * emulate_side_effects != 0 also leaves dL, cs*, dL_total, TdL_total, F, Pin as the reference's
 * single-threaded assembly leaves them. */
int lpmb_fd_stiffness(lpmb_ctx *ctx, int emulate_side_effects);
/* fill every stored block with a deterministic, symmetric, diagonally dominant test pattern
 * (bench utility: lets the SpMV be timed at sizes where no reference matrix exists yet) */
int lpmb_matrix_fill_test_pattern(lpmb_ctx *ctx);
/* y = K x on host DoF vectors (interleaved); for tests and the SpMV micro-benchmark */
int lpmb_spmv_host(lpmb_ctx *ctx, const double *x, double *y);
/* repeat y = K x `reps` times on device-resident vectors, returns mean milliseconds per SpMV
 * measured with CUDA events on the context stream.  variant: 0 = full-format SELL kernel, 2 = brick-blocked
 * symmetric kernel (needs lpmb_matrix_enable_bricks); 1 was an L2-mediated symmetric variant, removed (0.85x) */
int lpmb_spmv_bench(lpmb_ctx *ctx, int reps, int variant, double *ms_per_spmv);
/* Brick-blocked symmetric SpMV (lpmb_brick.cu): K is symmetric (stiffness.c:441-481 keeps one triangle), so
 * every block is streamed from HBM once and used for both of its contributions inside one CTA; about half the
 * bytes of the full format per CG iteration.  Needs radius, xyz_initial and the connectivity; applies to
 * axis-aligned simple-cubic 3-D lattices -- one GPU, or z-slabs of whole layers (then a collective call: either every
 * rank gets its bricks or none does) -- and returns LPMB_ERR_UNSUPPORTED (nothing enabled) otherwise.  Once on, lpmb_solve_cg* / lpmb_newton_iteration / lpmb_spmv_host use it; on = 0 releases it.
 * Particle numbering at the ABI is unchanged (the brick order is internal to the solve). */
int lpmb_matrix_enable_bricks(lpmb_ctx *ctx, int on);
/* bytes one brick SpMV moves (matrix + staging + vectors); 0 when bricks are off */
long long lpmb_spmv_bytes_bricks(lpmb_ctx *ctx);
/* algorithmic bytes one SpMV moves: nblk*(8 d^2+4) + 4 (N+1) + 16 d N */
long long lpmb_spmv_bytes(lpmb_ctx *ctx);
/* bytes the device format actually streams (SELL padding included) */
long long lpmb_spmv_bytes_stored(lpmb_ctx *ctx);

/* live profiling of the dominant kernel: when on, every CG SpMV launch is bracketed by CUDA events on
 * the context stream; lpmb_get_profile returns the accumulated device time and launch count */
int lpmb_set_profiling(lpmb_ctx *ctx, int on);
int lpmb_get_profile(lpmb_ctx *ctx, double *spmv_ms_total, long long *spmv_calls);

/* ---- linear solve -------------------------------------------------------------------------- */
/* DoF mask: 1 = free, 0 = constrained (dispBC_index[k] && fix_index[k]); NULL = all free. */
int lpmb_set_dof_mask(lpmb_ctx *ctx, const int *dispBC_index, const int *fix_index);
/* Unpreconditioned CG, x0 = 0, stop when ||r||^2 <= rel*||r0||^2 + abs or maxit iterations
 * (solver.c:217-222: rel=1e-8, abs=1e-12, maxit=n).  rhs/disp are host DoF vectors.
 * use_mask != 0 solves the BC-modified system without editing K.
 * Returns LPMB_ERR_NOTCONVERGED (disp still written) when maxit is hit. */
int lpmb_solve_cg(lpmb_ctx *ctx, const double *rhs, double *disp, double rel, double abs_tol, int maxit,
                  int use_mask, int *iterations);
/* same, entirely on device fields "residual" -> "disp"; optionally xyz += disp (solver.c:263-267) */
int lpmb_solve_cg_device(lpmb_ctx *ctx, double rel, double abs_tol, int maxit, int use_mask, int update_xyz,
                         int *iterations);

/* ---- constitutive path --------------------------------------------------------------------- */
int lpmb_calc_kntv(lpmb_ctx *ctx, const double *Ce, int ntype);
int lpmb_compute_dl(lpmb_ctx *ctx);
int lpmb_bond_force(lpmb_ctx *ctx, int plmode, int load_indicator);
/* The reference's per-particle law entry points (constitutive.h:15,17,18,20,21): computeBondForceElastic(ii) for plmode 6
 * (src/constitutive.c:228-283), computeBondForceIncrementalUpdating(ii) for plmode 4 (:167-225),
 * computeBondForceJ2mixedLinear3D(ii) for plmode 0 (:466-686), computeBondForceJ2energyReturnMap(ii, load_indicator) for
 * plmode 3 (:286-463), computeBondForceJ2nonlinearIso(ii) for plmode 5 (:689-863).  Same side effects as the reference:
 * geometry (and return-map) outputs of ii AND of its neighbours across intact bonds, F / Pin (and slot-[2] state,
 * J2_dlambda, dL_ave; plmode 0 / 5: stress_tensor[ii] := 0) of ii only; plmode 5 advances the slot-[0] plastic state of the
 * whole star in place and leaves the star members' trial forces in F; no computeStress, no switchStateV(2).  O(N) per
 * call (API completeness; the assembly and the whole-lattice laws never go through it).  plmode 1 =
 * computeBondForceCPMiehe(ii) (:866-1396) with its memo, the int field "state_v" (:946-959): star members flagged 1 REUSE
 * the increments ddLp / cp_dgy / cp_dA / cp_dA_single an earlier call left, the others are return-mapped and flagged;
 * lpmb_bond_force(1, .) leaves the memo all 1 like the reference's serial loop, the host zeroes it with lpmb_field_set.
 * Anything else: LPMB_ERR_UNSUPPORTED. */
int lpmb_bond_force_particle(lpmb_ctx *ctx, int plmode, int particle, int load_indicator);
int lpmb_switch_state(lpmb_ctx *ctx, int flag);
/* residual = dispBC_index*(Pex-Pin); returns ||residual||_2 and ||reaction||_2 (either may be NULL) */
int lpmb_update_rr(lpmb_ctx *ctx, double *norm_residual, double *norm_reaction);
/* returns the number of newly broken bonds in *broken; broken (i, neighbor) pairs are appended to
 * pairs[2*k], pairs[2*k+1] in the reference's logging order, up to max_pairs (may be NULL).
 * plmode 0 / 5 / 6 select what updateDamageGeneral dispatches to (constitutive.c:149-164); the two laws its
 * dispatcher keeps commented out (:155-156) are reached with the codes below.  LPMB_DAMAGE_PWISE_LOCAL counts and
 * reports detached PARTICLES (pairs[2k] = particle, pairs[2k+1] = -1), as the reference logs them (:1562). */
#define LPMB_DAMAGE_PWISE_LOCAL 100    /* updateDuctileDamagePwiseLocal,    src/constitutive.c:1529-1579 */
#define LPMB_DAMAGE_BWISE_NONLOCAL 101 /* updateDuctileDamageBwiseNonlocal, src/constitutive.c:1698-1753 */
int lpmb_update_damage(lpmb_ctx *ctx, int plmode, int *broken, int *pairs, int max_pairs);
/* computeStrain(), lpm_basic.c:127-249: per-particle weighted-least-squares strain tensor from the elastic bond
 * stretches dL over the initial bond directions (n x n LU with partial pivoting, n = 3(dim-1)); writes the field
 * strain_tensor[N][6] in the reference's component order (e11 e22 e33 e23 e13 e12). */
int lpmb_compute_strain(lpmb_ctx *ctx);
int lpmb_update_crack(lpmb_ctx *ctx);

/* ---- crystal plasticity (plmode 1) ---------------------------------------------------------- */
/* schmid_tensor[nslipSys][6] as slipSysDefine3D builds it (src/initialization.c:563-827); also sets the
 * parameter nslipSys that sizes the cp_* fields.  Parameters read by the law: cp_h0, cp_taus0, cp_tau00
 * (= cp_taus[0], cp_tau0[0]), cp_q, cp_eta, cp_p, cp_maxloop, dtime. */
int lpmb_set_schmid_tensor(lpmb_ctx *ctx, const double *schmid_tensor, int nslipSys);
/* computeCab(), src/constitutive.c:1864-1917: fills the field cp_Cab [N][nslipSys^2] */
int lpmb_compute_cab(lpmb_ctx *ctx);

/* ---- one Newton iteration, device resident (lpmc_project.c:426-464) ------------------------ */
/* switchStateV(0); masked CG solve; xyz += disp; computeBondForceGeneral(plmode); updateRR; norm */
int lpmb_newton_iteration(lpmb_ctx *ctx, int plmode, int load_indicator, double rel, double abs_tol, int maxit,
                          int *cg_iterations, double *norm_residual);

/* ---- device-resident driver helpers -------------------------------------------------------- */
/* dst := src for two fields of identical shape (the driver's xyz_temp/F_temp/Pex_temp copies,
 * lpmc_project.c:387-389) */
int lpmb_field_copy(lpmb_ctx *ctx, const char *dst, const char *src);
/* setDispBC / setForceBC for one table entry, src/boundary.c:12-70, applied to the resident
 * xyz / dispBC_index / Pex (axis is 'x', 'y' or 'z') */
int lpmb_apply_disp_bc(lpmb_ctx *ctx, int type, char axis, double step);
int lpmb_apply_force_bc(lpmb_ctx *ctx, int type, double step_x, double step_y, double step_z);

/* ---- binary snapshots (checkpoint / resume) ------------------------ */
/* The reference only has per-step TEXT dumps (data_handler.c:42-84) and in-memory roll-back state.  save: scalar
 * parameters + every field, in device layout, to one compact file.  load: into a context created with the same
 * (nparticle, dim, lattice, nneighbors, nconn_max); restores the state bit for bit, rebuilds the block pattern of K
 * and the DoF mask; the tangent VALUES are not stored -- call lpmb_fd_stiffness (every load step starts with it). */
int lpmb_snapshot_save(lpmb_ctx *ctx, const char *path);
int lpmb_snapshot_load(lpmb_ctx *ctx, const char *path);

/* ---- multi-GPU (one process per GPU; particle slabs = contiguous index ranges) ------------- */
/* 128-byte NCCL unique id; rank 0 creates it, the harness broadcasts it. */
int lpmb_dist_unique_id(void *id128);
int lpmb_dist_init(lpmb_ctx *ctx, const void *id128, int rank, int world);
/* Declare this rank's slab.  The context holds [ghost_lo | owned | ghost_hi] in global particle order
 * (a contiguous index range = z-slab + 4 ghost lattice layers on each inner side); owned = local
 * indices [own0, own1).  narrow_* = particle counts exchanged with rank-1 (lo) / rank+1 (hi) on every
 * CG iteration (2 layers: the reach of conn); wide exchanges fill all ghosts (recv counts = own0 and
 * N-own1), wide_send_* = what the neighbours expect from this rank.  After this call SpMV / dot
 * products / norms run on owned rows only and the CG all-reduces its scalars. */
int lpmb_dist_set_slab(lpmb_ctx *ctx, int own0, int own1, int narrow_recv_lo, int narrow_recv_hi, int narrow_send_lo,
                       int narrow_send_hi, int wide_send_lo, int wide_send_hi);
/* halo exchange of one named fp64 field (wide != 0: all ghosts, else the narrow CG halo) */
int lpmb_dist_exchange_field(lpmb_ctx *ctx, const char *name, int wide);
/* How the per-CG-iteration traffic of a slab run travels: 0 = single GPU, 1 = NCCL only, 2 = scalar all-reduces
 * through CUDA-IPC mapped peer memory over NVLink (halo through NCCL), 3 = scalars and the halo push of the search
 * direction through peer memory (brick SpMV enabled).  lpmb_dist_init maps the peers unless the environment
 * variable LPMB_NO_PEER is set or a rank cannot map another (then all ranks stay on NCCL together); param
 * "peer_comm" = 0 switches the fast path off at run time. */
int lpmb_dist_mode(lpmb_ctx *ctx);
/* The selection step of updateBrittleDamage (constitutive.c:1489-1520) on k candidates listed in the reference's scan order
 * (keys ascending): sorts keys / strains in place with the reference's shell sort when k > nbreak; the bonds to break are
 * [*first, k).  Pure host arithmetic (no device); lpmb_update_damage(ctx, 6, ...) uses it, in slab runs on the all-gathered
 * candidates of all ranks. */
int lpmb_brittle_select(int k, long long *keys, double *strains, int nbreak, int *first);
/* The multigrid levels of the fast mode (param cg_precond) for a full simple-cubic block of nx x ny x sum(owned) sites split
 * into z-slabs of owned[r] layers: pure host arithmetic, no device.  plan[8 l + k], k = 0: level distributed over the ranks
 * (else whole on every rank), 1 / 2: nx, ny of the level, 3: layers of THIS rank's block of it, 4: global z of its layer 0,
 * 5 / 6: its owned layers [oz0, oz1) inside the block, 7: layers of the whole level.  lrep = first replicated level (-1: none),
 * gat_off / gat_cnt [world] = the ranks' owned ranges of that level (elements per component).  nz_local0 / ghost_lo0: layers of
 * this rank's level-0 block and how many of them lie below the owned ones (world = 1: sum(owned), 0).  See lpmb_mg.cu. */
int lpmb_mg_slab_plan(int world, int rank, const long long *owned, int nx, int ny, int nz_local0, int ghost_lo0, int max_levels, int *plan,
                      long long *gat_off, long long *gat_cnt, int *nlev, int *lrep);

#ifdef __cplusplus
}
#endif
#endif
