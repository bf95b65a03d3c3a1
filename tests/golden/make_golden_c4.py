"""Generate tests/golden/c4_fcc_real.npz: BASELINE config 4 at its REAL size from the UNMODIFIED reference (oracle/_ref),
single thread.

examples/FCC_Al_R0.3_001_tension.c itself cannot be compiled against the reference's current src/ tree -- not for want of a
few globals (as the CT example): its own TU re-defines `damage_broken` as int** and `J2_beta_eq` as double* against
include/lpm.h:78,80 (double**) and calls updateStateJ2energyWei(), which exists nowhere; no shim TU can repair a type
conflict inside the example's translation unit.  So the example's set-up is replayed with the reference's CURRENT functions
(oracle/ref.py::setup_fcc follows its lines 62-229: FCC lattice 3, radius 0.3, box 0..10 -> 6 912 particles, Al elastic
constants, crystal plasticity plmode 1 with 24 slip systems, types 1-5) and its load loop (:285-330, :382-546 of the default
driver for the call order): nine displacement BCs, top layer pulled by -2e-3 per step, dtime 0.1, CG.

Recorded: Newton iteration counts, CG iteration counts of every solve, reaction-force norm per step, and after every load
step xyz, stress_tensor, cp_A and the number of active slip systems; after the LAST step also F, cp_gy, dLp (slot 0) and cp_Jact.
Run here (container with /root/reference; ~10 min):   python tests/golden/make_golden_c4.py [steps=3]
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import RefLPM  # noqa: E402

DBP = [(1, "z", 0.0), (2, "x", 0.0), (2, "z", 0.0), (3, "y", 0.0), (3, "z", 0.0), (4, "x", 0.0), (4, "y", 0.0), (4, "z", 0.0),
       (5, "z", -2.0e-3)]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    r = RefLPM.instance()
    r.threads(int(os.environ.get("LPMB_C4_SETUP_THREADS", "1")))
    t0 = time.time()
    r.setup_fcc()                       # box 0..10, radius 0.3 -> 6 912 particles (O(N^2) search + computeCab)
    r.threads(1)
    L = r.lib
    dropin = hasattr(L, "lpmc_dropin_last_cg_iterations")
    g = {"setup.xyz": r.get("xyz"), "setup.type": r.get("type"), "setup.nb_initial": r.get("nb_initial"),
         "sizes": np.array([r.N, int(r.get("nb_initial").sum()), int(r.get("nb_conn").sum()), int(r.get("K_pointer")[r.N, 1])])}
    print("set-up", round(time.time() - t0, 1), "s; N, bonds, conn blocks, nnz_upper =", g["sizes"].tolist(), flush=True)
    newton, cg_all, wall, law = [], [], [], []
    for step in range(1, steps + 1):
        ts = time.time()
        nr, nf = r.begin_step(DBP, [])
        tol, ni, cg = max(nr, nf), 0, []
        t_law = 0.0
        while nr > 1e-4 * tol and ni < 100:
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            cg.append(int(L.lpmc_dropin_last_cg_iterations() if dropin else L.lpmb_shim_last_itercount()))
            a = time.time()
            L.computeBondForceGeneral(1, 1)
            t_law += time.time() - a
            L.updateRR()
            nr = r.norms()[0]
            ni += 1
        jact = r.get_cp("cp_Jact")
        L.computeStrain()
        L.updateDamageGeneral(b"/dev/null", step, 1)
        L.updateCrack()
        L.switchStateV(1)
        newton.append(ni)
        cg_all.append(cg)
        wall.append(time.time() - ts)
        law.append(t_law)
        g[f"s{step}.xyz"] = r.get("xyz")
        g[f"s{step}.stress_tensor"] = r.get("stress_tensor")
        g[f"s{step}.cp_A"] = r.get_cp("cp_A")[:, 0]
        g[f"s{step}.active_systems"] = np.array([int(jact.sum())])
        if step == steps:     # the bond-sized arrays only once (fixture size)
            g[f"s{step}.F"] = r.get("F")
            g[f"s{step}.cp_gy"] = r.get_cp("cp_gy")[..., 0]
            g[f"s{step}.dLp"] = r.get("dLp")[..., 0]
            g[f"s{step}.cp_Jact_last"] = jact.astype(np.int8)
        g[f"s{step}.reaction_norm"] = np.array([r.norms()[1]])
        g[f"s{step}.cg_iterations"] = np.array(cg)
        print(f"step {step}: {ni} Newton iterations, CG {cg}, {wall[-1]:.1f} s (CP law {t_law:.1f} s), active systems "
              f"{int(jact.sum())}, max cp_A {g[f's{step}.cp_A'].max():.3e}", flush=True)
    g["newton_counts"] = np.array(newton)
    g["wall_s"] = np.array(wall)
    g["law_s"] = np.array(law)        # computeBondForceGeneral(1) per load step (serial in the reference, constitutive.c:114-117)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "c4_fcc_real.npz"))
    np.savez_compressed(out, **g)
    print("wrote", out, round(out.stat().st_size / 1e6, 2), "MB")


if __name__ == "__main__":
    main()
