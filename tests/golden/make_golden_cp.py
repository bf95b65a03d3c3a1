"""Generate tests/golden/fcc_cp.npz from the UNMODIFIED reference (oracle/_ref): crystal plasticity (plmode 1,
computeBondForceCPMiehe + computeCab, src/constitutive.c:866-1396,1864-1917) on a 256-particle FCC block with the
material / BCs of examples/FCC_Al_R0.3_001_tension.c.  Run: python tests/golden/make_golden_cp.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import RefLPM  # noqa: E402

DBP = [(1, "z", 0.0), (2, "x", 0.0), (2, "z", 0.0), (3, "y", 0.0), (3, "z", 0.0), (4, "x", 0.0), (4, "y", 0.0), (4, "z", 0.0),
       (5, "z", -2.0e-3)]
BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "J2_stresseq", "J2_stressm", "J2_triaxiality", "xyz", "Pin", "pl_flag")
CP = ("cp_gy", "cp_A", "cp_A_single", "cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single")


def state(r, prefix, g):
    for n in BOND + PART:
        g[f"{prefix}.{n}"] = r.get(n)
    g[f"{prefix}.dLp"] = r.get("dLp")
    for n in CP:
        g[f"{prefix}.{n}"] = r.get_cp(n)


# BCC variant (LPMB_CP_LATTICE=4 -> tests/golden/bcc_cp.npz): 8 + 6 neighbours, 41 conn, 24 slip systems; the x- / y-line
# types of the FCC example are empty on this lattice, so both end layers are clamped in x and y instead
DBP_BCC = [(1, "x", 0.0), (1, "y", 0.0), (1, "z", 0.0), (5, "x", 0.0), (5, "y", 0.0), (5, "z", -2.0e-3)]


def main():
    import os
    global DBP
    lattice = int(os.environ.get("LPMB_CP_LATTICE", 3))
    if lattice == 4:
        DBP = DBP_BCC
    r = RefLPM.instance()
    r.threads(1)
    r.setup_fcc(box=(0, 3.5, 0, 3.5, 0, 3.5), lattice=lattice)
    L = r.lib
    g = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "nb_conn", "K_pointer", "type", "distance_initial", "csx_initial",
              "csy_initial", "csz_initial", "Kn", "Tv", "Ce", "distance", "dL", "csx", "csy", "csz", "dL_total", "TdL_total"):
        g[f"setup.{n}"] = r.get(n)
    g["setup.schmid_tensor"] = r.get_cp("schmid_tensor")
    g["setup.cp_Cab"] = r.get_cp("cp_Cab")
    g["setup.cp_gy"] = r.get_cp("cp_gy")
    names = ["radius", "particle_volume", "dtime", "cp_h0", "cp_p", "cp_q", "cp_eta", "cp_maxloop", "neighbor1_cutoff", "neighbor2_cutoff"]
    g["param_names"] = np.array(names + ["cp_tau00", "cp_taus0", "nslipSys"])
    g["params"] = np.array([r.gd(n) for n in names] + [r.darr("cp_tau0", 3)[0], r.darr("cp_taus", 3)[0], float(r.gi("nslipSys"))])
    counts = []
    for step in (1, 2):
        s = f"s{step}"
        nr, nf = r.begin_step(DBP, [])
        g[f"{s}.bc.dispBC_index"] = r.get("dispBC_index")
        state(r, f"{s}.pred", g)
        tol = max(nr, nf)
        ni = 0
        while nr > 1e-4 * tol and ni < 100:
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            if ni < 3:
                g[f"{s}.n{ni}.xyz"] = r.get("xyz")
            L.computeBondForceGeneral(1, 1)
            if ni < 3:
                state(r, f"{s}.n{ni}.bf", g)
            L.updateRR()
            nr = r.norms()[0]
            ni += 1
        counts.append(ni)
        L.computeStrain()
        L.updateDamageGeneral(b"/dev/null", step, 1)
        L.updateCrack()
        L.switchStateV(1)
        state(r, f"{s}.end", g)
    g["newton_counts"] = np.array(counts)
    g["dbp_type"] = np.array([t for (t, _, _) in DBP])
    g["dbp_axis"] = np.array([a for (_, a, _) in DBP])
    g["dbp_step"] = np.array([v for (_, _, v) in DBP])
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / ("bcc_cp.npz" if lattice == 4 else "fcc_cp.npz")))
    np.savez_compressed(out, **g)
    print("wrote", out, round(out.stat().st_size / 1e6, 2), "MB; newton iterations per step:", counts)


if __name__ == "__main__":
    main()
