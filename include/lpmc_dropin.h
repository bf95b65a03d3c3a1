/*
 * lpmc_dropin.h -- the reference-named entry points of the hot path, implemented on the B200.
 *
 * liblpmc_dropin (lpm-c_b200/csrc/dropin.c) defines, with C linkage and the reference's exact
 * signatures, the functions of the reference's three hot-path headers so that the reference's own
 * drivers (src/lpmc_project.c, examples/ *.c) and its remaining host translation units (boundary.c,
 * initialization.c, neighbor.c, lpm_basic.c, data_handler.c) link UNCHANGED -- simply leave
 * src/stiffness.c, src/solver.c and src/constitutive.c out of the link and add
 * -llpmc_dropin -llpmb200.  Each one marshals the reference's jagged process
 * globals (include/lpm.h:55-81 of the reference) to the C ABI of include/lpmb200.h and back.
 *
 *   reference header : line    entry point                         GPU path behind it
 *   stiffness.h:5              void calcKnTv()                     lpmb_calc_kntv
 *   stiffness.h:6              void updateRR()                     lpmb_update_rr (+ reaction_force compaction)
 *   stiffness.h:7-8            void calcStiffness{2,3}DFiniteDifference(int plmode)
 *                                                                  lpmb_fd_stiffness(emulate_side_effects=1)
 *                                                                  + lpmb_matrix_to_upper_csr -> K_global/IK/JK
 *   solver.h:6                 void solverCG()                     lpmb_matrix_from_upper_csr(K_global) + lpmb_solve_cg
 *   solver.h:5                 void solverPARDISO()                same iterative solver at 1e-12 relative residual
 *   constitutive.h:9           void switchStateV(int)              lpmb_switch_state
 *   constitutive.h:14          void computeBondForceGeneral(int plmode, int t)   lpmb_bond_force
 *   constitutive.h:22-27       int updateDamageGeneral / updateBrittleDamage / updateDuctileDamagePwiseNonlocal
 *                                                                  lpmb_update_damage (+ the broken-bond log file)
 *   constitutive.h:29          void updateCrack()                  lpmb_update_crack
 *   constitutive.h:11          void computeCab()                   lpmb_set_schmid_tensor + lpmb_compute_cab
 *   constitutive.h:25          int updateDuctileDamageBwiseLocal(const char *, int)   lpmb_update_damage (plmode 5)
 *   constitutive.h:24,26       int updateDuctileDamagePwiseLocal / updateDuctileDamageBwiseNonlocal(const char *, int)
 *                              (commented out in the reference's dispatcher, constitutive.c:155-156)
 *                                                                  lpmb_update_damage (LPMB_DAMAGE_PWISE_LOCAL / _BWISE_NONLOCAL)
 *   constitutive.h:15,17,20    void computeBondForceElastic(int) / computeBondForceJ2mixedLinear3D(int) /
 *                              computeBondForceIncrementalUpdating(int)                lpmb_bond_force_particle (6 / 0 / 4)
 *   constitutive.h:18,21       void computeBondForceJ2nonlinearIso(int) / computeBondForceJ2energyReturnMap(int, int)
 *                                                                  lpmb_bond_force_particle (5 / 3): one call of the
 *                              reference's serial loop -- plmode 5 advances slot [0] of ii's whole star in place
 *   constitutive.h:19          void computeBondForceCPMiehe(int)   lpmb_bond_force_particle (1): honours the host-visible memo
 *                              state_v like the reference (constitutive.c:946-959: flagged star members REUSE the increments an
 *                              earlier call left, the others are return-mapped and flagged); computeBondForceGeneral(1, .)
 *                              leaves state_v all 1, as the reference's serial loop does
 *
 * State ownership: the arrays these functions write (plastic state slots, damage_broken / damage_D / damage_w, nb,
 * bond forces ...) are uploaded once, before the first force evaluation -- so initial cracks set by the driver are
 * honoured -- and are device-authoritative afterwards: every call downloads what it wrote, but a driver that pokes
 * those arrays on the host BETWEEN calls is not seen (no shipped driver does).  Arrays only the host writes
 * (xyz, xyz_temp, F_temp, Pex, dispBC_index, fix_index, residual, K_global, type, Ce, parameters) are uploaded on
 * entry of the call that reads them.
 *
 * A second set-up in the same process (initMatrices() + the neighbour searches run again: the reference allocates new
 * arrays) is recognised by the addresses of neighbors / xyz_initial / conn / dLp and the sizes; the device context is then
 * rebuilt from the new host arrays on the next call.
 *
 * solverPARDISO() (solver.c:3-92; selected by no shipped driver): no sparse direct factorisation on the GPU path.  The
 * call runs the same CG to ||r|| <= 1e-12 ||r0|| (at most dim*N iterations), prints one notice on stderr the first time,
 * and exits with status 3 if that residual is not reached -- the reference's PARDISO path also exits on failure
 * (solver.c:50-84) rather than applying an unconverged displacement.
 *
 * Declarations use empty parameter lists exactly like the reference's headers (the default driver
 * even calls updateRR(ni++), lpmc_project.c:462 -- harmless under the SysV x86-64 ABI).
 *
 * Environment:
 *   LPMB_DEVICE=<n>            CUDA device (default 0)
 *   LPMB_DROPIN_BRICKS=1|0     force the brick-blocked symmetric CG SpMV on / off (default: on from 2^18 particles
 *                              on axis-aligned simple-cubic 3-D lattices; see lpmb_matrix_enable_bricks)
 *   LPMB_DROPIN_FAST=1         opt-in fast mode of solverCG(): CG preconditioned with the matrix-free multigrid V-cycle
 *                              (param cg_precond).  NOT the parity path (another Krylov sequence, same
 *                              stop rule); works on the re-imported K_global as well;
 *                              full simple-cubic blocks only, any other lattice makes solverCG() exit with the library's message
 *   LPMB_DROPIN_PROFILE=1      wall time per entry point on stderr at exit
 *   LPMB_DROPIN_DEVICE_BC=1    solverCG() keeps the tangent on the device and applies the displacement
 *                              BCs as a DoF mask instead of re-uploading the host-edited K_global
 *                              (identical iterates, see tests/test_solver_gpu.py; default off = strict)
 */
#ifndef LPMC_DROPIN_H
#define LPMC_DROPIN_H

#ifdef __cplusplus
extern "C" {
#endif

/* stiffness.h */
void calcKnTv();
void updateRR();
void calcStiffness2DFiniteDifference(int plmode);
void calcStiffness3DFiniteDifference(int plmode);
/* solver.h */
void solverPARDISO();
void solverCG();
/* constitutive.h */
void switchStateV(int conv_flag);
void computeCab();
void computeBondForceGeneral(int plmode, int temp);
void computeBondForceElastic(int i);
void computeBondForceJ2mixedLinear3D(int ii);
void computeBondForceJ2nonlinearIso(int ii);
void computeBondForceCPMiehe(int ii);
void computeBondForceIncrementalUpdating(int ii);
void computeBondForceJ2energyReturnMap(int ii, int load_indicator);
int updateDamageGeneral(const char *dataName, int tstep, int plmode);
int updateBrittleDamage(const char *dataName, int tstep, int nbreak);
int updateDuctileDamageBwiseLocal(const char *dataName, int tstep);
int updateDuctileDamagePwiseLocal(const char *dataName, int tstep);
int updateDuctileDamageBwiseNonlocal(const char *dataName, int tstep);
int updateDuctileDamagePwiseNonlocal(const char *dataName, int tstep);
void updateCrack();

/* Tell the layer that the host copies of device-authoritative state arrays (damage_*, plastic state slots, F, dL,
 * cs*, nb, J2_dlambda, J2_triaxiality ...) were edited by the caller since the last entry point returned: the next
 * entry point uploads them again.  Not needed by any shipped driver (they only edit those arrays before the first
 * force evaluation); test generators that poke state between calls use it. */
void lpmc_dropin_invalidate_state(void);

/* CG iterations of the last solverCG() / solverPARDISO() call (the reference only prints them, solver.c:254) */
int lpmc_dropin_last_cg_iterations(void);

/* release the device context (optional; the process exit does it too) */
void lpmc_dropin_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif
