"""Generate tests/golden/sc6_j2iso.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so), single thread.

plmode 5 = computeBondForceJ2nonlinearIso (src/constitutive.c:689-863) + updateDuctileDamageBwiseLocal
(:1607-1695).  No shipped driver selects it.  The law updates slot [0] of the plastic
state IN PLACE for particle ii and each of its intact neighbours on every call, so over one
computeBondForceGeneral(5, .) every particle is returned-mapped 1 + nb times in the order of the serial particle
loop, the bond forces of ii are formed from whatever its neighbours hold at that moment, F[i] is finally left
holding the TRIAL forces of the last call that touched i, and switchStateV(2) then overwrites slot [0] with the
never-written slot [2].  The fixture pins exactly that single-threaded behaviour: for every recorded call the
complete input state and every output.  6^3 block of sc6_j2.npz, top layer loaded -9000 (=> beyond SY(0) = 620 MPa),
J2_C = 2000 so the back stress moves.
Run here (container with /root/reference):   python tests/golden/make_golden_j2iso.py
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
        "J2_alpha", "J2_beta_eq", "xyz", "Pin", "pl_flag", "nb", "damage_local")


def state(r, prefix, out):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.dLp"] = r.get("dLp")
    out[f"{prefix}.J2_beta"] = r.get("J2_beta")
    out[f"{prefix}.damage_D"] = r.get("damage_D")


LOAD2 = float(os.environ.get("LPMB_J2ISO_LOAD2", -3000.0))


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5, plmode=5, damagec_A=1.5, damage_threshold=0.02)
    r.sd("J2_C", 2000.0)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    assert N == 216, N
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "type", "distance_initial", "csx_initial", "csy_initial",
              "csz_initial", "Kn", "Tv", "Ce", "sigmay"):
        g[f"setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "J2_H", "J2_xi", "J2_C", "damage_L", "damage_threshold", "damagec_A", "neighbor1_cutoff",
             "neighbor2_cutoff"]
    g["params"] = np.array([r.gd(n) for n in names])
    g["param_names"] = np.array(names)
    counts = []
    for step in (1, 2, 3):
        s = f"s{step}"
        r.set_d2("xyz_temp", r.d2("xyz", N, 3))
        r.set_d2("F_temp", r.d2("F", N, nn))
        r.set_d1("Pex_temp", r.d1("Pex", dim * N))
        L.calcStiffness3DFiniteDifference(6)
        # the bottom layer is clamped in all three directions: the law's bond forces are not self-equilibrated, and
        # the default driver's z-only support would leave the rigid-body modes to the CG
        d_arr = (DispBCPara * 3)(DispBCPara(1, b"x", 0.0), DispBCPara(1, b"y", 0.0), DispBCPara(1, b"z", 0.0))
        f_arr = (ForceBCPara * 1)(ForceBCPara(2, b"x", 0.0, b"y", 0.0, b"z", -9000.0 if step == 1 else LOAD2))
        L.setDispBC(3, d_arr)
        L.setForceBC(1, f_arr)
        L.computeBondForceGeneral(4, 1)
        L.updateRR()
        nr, nf = r.norms()
        tol = max(nr, nf)
        ni = 0
        while nr > 1e-4 * tol and ni < (1 if step == 2 else 2):
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            state(r, f"{s}.n{ni}.pre", g)
            L.computeBondForceGeneral(5, 1)
            state(r, f"{s}.n{ni}.bf", g)
            L.updateRR()
            nr = r.norms()[0]
            ni += 1
        counts.append(ni)
        # J2_dlambda holds the LAST return map of the call (0: the earlier ones already consumed the excess), so the
        # local damage law never sees plastic flow here; feed it a synthetic multiplier field instead
        # (step 2 only; centred in a lattice cell so that bonds break but no particle loses all of them -- a fully
        # detached particle makes the law divide by nb = 0)
        if step == 2:
            x0 = g["setup.xyz"]
            r.put("J2_dlambda", 0.03 * np.exp(-4.0 * ((x0 - np.array([1.25, 1.25, 1.25])) ** 2).sum(axis=1)))
        state(r, f"{s}.dam.pre", g)
        broken = L.updateDamageGeneral(b"/dev/null", step, 5)
        g[f"{s}.dam.broken"] = np.array([broken])
        state(r, f"{s}.dam", g)
        L.updateCrack()
        L.switchStateV(1)
    g["newton_counts"] = np.array(counts)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_j2iso.npz"))
    np.savez_compressed(out, **g)
    for k in ("s1.n0", "s1.n1", "s2.n0", "s3.n0", "s3.n1"):
        b = g[f"{k}.bf.J2_dlambda"]
        print(k, "finite", bool(np.isfinite(g[f"{k}.bf.F"]).all() and np.isfinite(g[f"{k}.bf.Pin"]).all()), "dlambda>0", int((b > 0).sum()),
              "max", b.max(), "|Pin|", np.abs(g[f"{k}.bf.Pin"]).max(), "dLp0 max", np.abs(g[f"{k}.bf.dLp"][..., 0]).max(),
              "nb min", g[f"{k}.pre.nb"].min())
    print("broken per step", [int(g[f"s{k}.dam.broken"][0]) for k in (1, 2, 3)], "damage_local max", g["s3.dam.damage_local"].max(),
          "nb min", g["s3.dam.nb"].min())
    print("wrote", out, out.stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    main()
