"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The reference ships no tests or fixtures of its own, so these vectors are
the pinned outputs of the reference's own functions, single-threaded, strict IEEE
(-ffp-contract=off), driven through oracle/ref.py exactly in the order of the default driver's
loop (src/lpmc_project.c:382-546).

sc6_j2.npz : 6x6x6 simple-cubic block (216 particles), default material of lpmc_project.c
             (E=146e3, nu=0.3, sigma_y=200, H=38.714e3, isotropic), bottom z-layer fixed in z,
             top z-layer loaded -2000 in z per step (=> plastic from step 1: 2000/9 > 200).
             Two load steps; per Newton iteration the inputs and outputs of every hot-path call.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import RefLPM  # noqa: E402

BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality",
        "J2_alpha", "damage_nonlocal", "xyz", "Pin", "pl_flag", "residual")


def state(r: RefLPM, prefix: str, out: dict):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.dLp"] = r.get("dLp")
    out[f"{prefix}.J2_beta"] = r.get("J2_beta")
    out[f"{prefix}.damage_D"] = r.get("damage_D")


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    assert N == 216, N
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "nb_conn", "K_pointer", "type", "distance_initial",
              "csx_initial", "csy_initial", "csz_initial", "Kn", "Tv", "Ce", "KnTve", "sigmay", "distance", "dL",
              "dL_total", "TdL_total", "csx", "csy", "csz"):
        g[f"setup.{n}"] = r.get(n)
    g["params"] = np.array([r.gd("radius"), r.gd("particle_volume"), r.gd("J2_H"), r.gd("J2_xi"), r.gd("damage_L"),
                            r.gd("damage_threshold"), r.gd("damagec_A"), r.gd("neighbor1_cutoff"),
                            r.gd("neighbor2_cutoff")])
    g["param_names"] = np.array(["radius", "particle_volume", "J2_H", "J2_xi", "damage_L", "damage_threshold",
                                 "damagec_A", "neighbor1_cutoff", "neighbor2_cutoff"])
    dbp, fbp = [(1, "z", 0.0)], [(2, 0.0, 0.0, -2000.0)]
    newton_counts = []
    for step in (1, 2):
        s = f"s{step}"
        g[f"{s}.pre.xyz"] = r.get("xyz")
        g[f"{s}.pre.F"] = r.get("F")
        g[f"{s}.pre.dLp"] = r.get("dLp")
        # --- lpmc_project.c:387-414 with snapshots between the calls
        r.set_d2("xyz_temp", r.d2("xyz", N, 3))
        r.set_d2("F_temp", r.d2("F", N, nn))
        r.set_d1("Pex_temp", r.d1("Pex", dim * N))
        L.calcStiffness3DFiniteDifference(6)
        for n in ("K_global", "IK", "JK"):
            g[f"{s}.fd.{n}"] = r.get(n)
        state(r, f"{s}.fd", g)  # side effects of the assembly
        nr, nf = None, None
        from oracle.ref import DispBCPara, ForceBCPara
        d_arr = (DispBCPara * 1)(DispBCPara(1, b"z", 0.0))
        f_arr = (ForceBCPara * 1)(ForceBCPara(2, b"x", 0.0, b"y", 0.0, b"z", -2000.0))
        L.setDispBC(1, d_arr)
        L.setForceBC(1, f_arr)
        g[f"{s}.bc.xyz"] = r.get("xyz")
        g[f"{s}.bc.Pex"] = r.get("Pex")
        g[f"{s}.bc.dispBC_index"] = r.get("dispBC_index")
        g[f"{s}.bc.fix_index"] = r.get("fix_index")
        L.computeBondForceGeneral(4, 1)
        state(r, f"{s}.pred", g)
        g[f"{s}.pred.ddL"] = r.get("ddL")
        g[f"{s}.pred.ddL_total"] = r.get("ddL_total")
        g[f"{s}.pred.TddL_total"] = r.get("TddL_total")
        L.updateRR()
        nr, nf = r.norms()
        g[f"{s}.rr.residual"] = r.get("residual")
        g[f"{s}.rr.reaction_force"] = r.get("reaction_force")
        g[f"{s}.rr.norms"] = np.array([nr, nf])
        tol = max(nr, nf)
        ni = 0
        while nr > 1e-4 * tol and ni < 100:
            t = f"{s}.n{ni}"
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            if ni < 3:
                g[f"{t}.K_bc"] = r.get("K_global")
                g[f"{t}.rhs"] = r.get("residual")
            L.solverCG()
            if ni < 3:
                g[f"{t}.disp"] = r.get("disp")
                g[f"{t}.cg_iters"] = np.array([L.lpmb_shim_last_itercount()])
                g[f"{t}.xyz"] = r.get("xyz")
            L.computeBondForceGeneral(r.gi("plmode"), 1)
            if ni < 3:
                state(r, f"{t}.bf", g)
            L.updateRR()
            nr = r.norms()[0]
            if ni < 3:
                g[f"{t}.residual"] = r.get("residual")
                g[f"{t}.norm_residual"] = np.array([nr])
            ni += 1
        newton_counts.append(ni)
        g[f"{s}.strain.dL"] = r.get("dL")
        L.computeStrain()
        g[f"{s}.strain.strain_tensor"] = r.get("strain_tensor")
        broken = L.updateDamageGeneral(b"/dev/null", step, r.gi("plmode"))
        g[f"{s}.dam.broken"] = np.array([broken])
        state(r, f"{s}.dam", g)
        L.updateCrack()
        state(r, f"{s}.crack", g)
        g[f"{s}.crack.nb"] = r.get("nb")
        g[f"{s}.crack.damage_visual"] = r.get("damage_visual")
        L.switchStateV(1)
        g[f"{s}.commit.dLp"] = r.get("dLp")
        g[f"{s}.commit.J2_alpha"] = r.get("J2_alpha")
    g["newton_counts"] = np.array(newton_counts)
    import os
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_j2.npz"))
    np.savez_compressed(out, **g)
    print("wrote", out, out.stat().st_size / 1e6, "MB; newton iterations per step:", newton_counts)


if __name__ == "__main__":
    main()
