"""Generate tests/golden/sc6_j2energy.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so).

plmode 3 = computeBondForceJ2energyReturnMap (src/constitutive.c:286-463): J2 plasticity with the
distortional-energy return map, scalar back stress, plastic multiplier by bisection.  No shipped driver
selects it, so the fixture replays the default driver's load-step loop
(src/lpmc_project.c:382-546) on the 6^3 block of sc6_j2.npz with plmode = 3 and records, for every call of
computeBondForceGeneral(3, t), the complete input state and every output.  Scenarios:
  s1.n*   first load step (top layer loaded -2000 in z => plastic), t = +1, mixed hardening J2_xi = 0.3
  s2.n*   second step after the state was committed, t = -1 (reversed load indicator), three bonds broken and
          updateCrack() run first, so nb[i] < nb_initial[i] (the law loops over the first nb[i] slots)
Run here (container with /root/reference):   python tests/golden/make_golden_j2e.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import DispBCPara, ForceBCPara, RefLPM  # noqa: E402

BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality",
        "J2_alpha", "J2_beta_eq", "xyz", "Pin", "pl_flag", "nb")


def state(r, prefix, out):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.dLp"] = r.get("dLp")
    out[f"{prefix}.damage_D"] = r.get("damage_D")


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5, plmode=3, J2_xi=0.3)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    assert N == 216, N
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "type", "distance_initial", "csx_initial", "csy_initial",
              "csz_initial", "Kn", "Tv", "Ce", "sigmay"):
        g[f"setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "J2_H", "J2_xi", "damage_L", "damage_threshold", "damagec_A", "neighbor1_cutoff",
             "neighbor2_cutoff"]
    g["params"] = np.array([r.gd(n) for n in names])
    g["param_names"] = np.array(names)
    counts = []
    for step, t in ((1, 1), (2, -1)):
        s = f"s{step}"
        if step == 2:
            # break three bonds symmetrically, recount nb (updateCrack, constitutive.c:1399-1434)
            nbr, br = r.get("neighbors"), r.get("damage_broken")
            w = r.get("damage_w")
            for i, j in ((100, 0), (57, 5), (130, 11)):
                k = int(nbr[i, j])
                jj = int(np.where(nbr[k] == i)[0][0])
                br[i, j] = br[k, jj] = 0.0
                w[i, j] = w[k, jj] = 0.0
            r.put("damage_broken", br)
            r.put("damage_w", w)
            if hasattr(L, "lpmc_dropin_invalidate_state"):   # drop-in replay: announce the HOST edit (include/lpmc_dropin.h)
                L.lpmc_dropin_invalidate_state()
            L.updateCrack()
        r.set_d2("xyz_temp", r.d2("xyz", N, 3))
        r.set_d2("F_temp", r.d2("F", N, nn))
        r.set_d1("Pex_temp", r.d1("Pex", dim * N))
        L.calcStiffness3DFiniteDifference(6)
        d_arr = (DispBCPara * 1)(DispBCPara(1, b"z", 0.0))
        f_arr = (ForceBCPara * 1)(ForceBCPara(2, b"x", 0.0, b"y", 0.0, b"z", -2000.0 * t))
        L.setDispBC(1, d_arr)
        L.setForceBC(1, f_arr)
        L.computeBondForceGeneral(4, t)
        L.updateRR()
        nr, nf = r.norms()
        tol = max(nr, nf)
        ni = 0
        while nr > 1e-4 * tol and ni < 3:
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            state(r, f"{s}.n{ni}.pre", g)
            L.computeBondForceGeneral(3, t)
            state(r, f"{s}.n{ni}.bf", g)
            L.updateRR()
            nr = r.norms()[0]
            g[f"{s}.n{ni}.norm_residual"] = np.array([nr])
            ni += 1
        counts.append(ni)
        L.switchStateV(1)
    g["newton_counts"] = np.array(counts)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_j2energy.npz"))
    np.savez_compressed(out, **g)
    pl = [int(g[f"s{k}.n0.bf.pl_flag"].sum()) for k in (1, 2)]
    print("wrote", out, out.stat().st_size / 1e6, "MB; newton iterations per step:", counts, "plastic particles:", pl,
          "max dlambda", [float(g[f"s{k}.n0.bf.J2_dlambda"].max()) for k in (1, 2)])


if __name__ == "__main__":
    main()
