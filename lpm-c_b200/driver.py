"""Device-resident replay of the reference driver's load-step loop (src/lpmc_project.c:382-546).

Host-side control flow only: every array stays in HBM and every numerical step is one C-ABI call
(include/lpmb200.h).  This is the "scaled synthetic driver" of SURVEY section 8(f): the shipped drivers
are literal-edited main()s whose O(N^2) set-up and per-step text dumps cannot reach 10M particles;
this one sets the same quantities through the ABI and keeps the same order of operations:

    xyz_temp, F_temp, Pex_temp := xyz, F, Pex            :387-389
    calcStiffness{2,3}DFiniteDifference(6)                :393-396
    setDispBC, setForceBC                                 :402-403
    computeBondForceGeneral(4, t)  (predictor)            :405
  label_broken_bond:
    updateRR, norms, tolerance multiplier                 :409-414
    while ||res|| > TOLITER*max(||res||0, ||reaction||0) and ni < MAXITER:   :424
        switchStateV(0); BC; solverCG; computeBondForceGeneral(plmode); updateRR     :428-463
    computeStrain (output record; compute_strain=True)     :466
    updateDamageGeneral; updateCrack; switchStateV(1)     :469-471
    if broken: calcStiffness...(6); goto label_broken_bond  :525-541
"""
from __future__ import annotations

from dataclasses import dataclass, field

TOLITER = 1e-4   # include/lpm.h:41
MAXITER = 100    # include/lpm.h:39


@dataclass
class StepLog:
    newton_iterations: int = 0
    cg_iterations: list = field(default_factory=list)
    residual_norms: list = field(default_factory=list)
    broken: int = 0
    reassemblies: int = 0


def load_step(ctx, plmode: int, disp_bc, force_bc, load_indicator: int = 1, rel=1e-8, abs_tol=1e-12,
              emulate_side_effects: bool = True, max_iter: int = MAXITER, compute_strain: bool = False) -> StepLog:
    """One load step.  disp_bc = [(type, axis, step)], force_bc = [(type, sx, sy, sz)] as dBP / fBP rows."""
    log = StepLog()
    ctx.copy_field("xyz_temp", "xyz")
    ctx.copy_field("F_temp", "F")
    ctx.copy_field("Pex_temp", "Pex")
    ctx.fd_stiffness(emulate_side_effects)
    for (t, axis, step) in disp_bc:
        ctx.apply_disp_bc(t, axis, step)
    for (t, sx, sy, sz) in force_bc:
        ctx.apply_force_bc(t, sx, sy, sz)
    ctx.bond_force(4, load_indicator)
    while True:
        nr, nf = ctx.update_rr()
        tol_mult = max(nr, nf)
        ni = 0
        while nr > TOLITER * tol_mult and ni < max_iter:
            it, nr = ctx.newton_iteration(plmode, load_indicator, rel, abs_tol)
            log.cg_iterations.append(it)
            log.residual_norms.append(nr)
            ni += 1
        log.newton_iterations += ni
        if compute_strain:
            ctx.compute_strain()
        broken, _ = ctx.update_damage(plmode)
        ctx.update_crack()
        ctx.switch_state(1)
        log.broken += broken
        if broken <= 0:
            break
        ctx.fd_stiffness(emulate_side_effects)
        log.reassemblies += 1
    return log
