"""Generate tests/golden/sc6_particle.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so), single thread.

The reference's PER-PARTICLE law entry points (include/constitutive.h:15,17,20), called directly the way
stiffness.c calls computeBondForceElastic(ii) -- outside the computeBondForceGeneral dispatcher:
    computeBondForceIncrementalUpdating(ii)   src/constitutive.c:167-225   (plmode 4)
    computeBondForceJ2mixedLinear3D(ii)       src/constitutive.c:466-686   (plmode 0)
    computeBondForceElastic(ii)               src/constitutive.c:228-283   (plmode 6)
Each call rewrites the geometry (and return-map) outputs of ii AND of its neighbours across intact bonds, and the
force outputs of ii only.  Case: the plastic 6^3 block of sc6_j2.npz with one bond broken before the first call
(as defineCrack does, initialization.c:1097-1118).  In load step 1 the calls are made at the points of the driver
loop where the dispatcher would run the same law (predictor after the BCs moved; J2 and elastic law after the first CG
solve of load step 1 and of load step 2, i.e. without and with plastic history); for every phase the complete state before
the first call and, after every call, all arrays the law may write.
Run here (container with /root/reference):   python tests/golden/make_golden_particle.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import DispBCPara, ForceBCPara, RefLPM  # noqa: E402

BOND = ("dL", "dL_ave", "ddLp", "ddL", "csx", "csy", "csz", "F", "F_temp", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "ddL_total", "TddL_total", "stress_tensor", "J2_dlambda", "J2_alpha", "xyz", "xyz_temp", "Pin", "pl_flag", "nb")
WRITES = {
    4: ("ddL", "ddL_total", "TddL_total", "F", "Pin"),
    6: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin"),
    0: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "stress_tensor", "J2_dlambda", "dLp2",
        "J2_beta2", "J2_alpha2"),   # slot [2] of the three-slot state arrays (slots [0], [1] are only read)
}
PARTICLES = (86, 0, 129, 215, 51)   # 86 and 51 share the broken bond; corner; interior; loaded top layer


def state(r, prefix, out):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.dLp"] = r.get("dLp")
    out[f"{prefix}.J2_beta"] = r.get("J2_beta")


def written(r, n):
    return r.get(n[:-1])[..., 2] if n.endswith("2") else r.get(n)


def phase(r, g, tag, law, fn):
    """pre-state, then one call per particle with the written arrays recorded after each"""
    if hasattr(r.lib, "lpmc_dropin_invalidate_state"):   # replay through the GPU drop-in layer
        r.lib.lpmc_dropin_invalidate_state()
    state(r, f"{tag}.pre", g)
    for k, ii in enumerate(PARTICLES):
        fn(ii)
        for n in WRITES[law]:
            g[f"{tag}.c{k}.{n}"] = written(r, n)
    g[f"{tag}.particles"] = np.array(PARTICLES)


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    assert N == 216, N
    # one bond broken in both directions, consistently (damage_broken, damage_w, damage_D, nb): defineCrack's writes
    nbr = r.get("neighbors")
    i0, s0 = 86, 3
    j0 = int(nbr[i0, s0])
    s1 = list(nbr[j0]).index(i0)
    assert j0 == PARTICLES[4], j0
    b, w, D, nb = r.get("damage_broken"), r.get("damage_w"), r.get("damage_D"), r.get("nb")
    for (i, s) in ((i0, s0), (j0, s1)):
        b[i, s] = 0.0
        w[i, s] = 0.0
        D[i, s, :] = 1.0
        nb[i] -= 1
    r.put("damage_broken", b); r.put("damage_w", w); r.put("damage_D", D); r.put("nb", nb)
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "type", "distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn",
              "Tv", "Ce", "sigmay"):
        g[f"setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "J2_H", "J2_xi", "damage_L", "damage_threshold", "damagec_A", "neighbor1_cutoff", "neighbor2_cutoff"]
    g["params"] = np.array([r.gd(n) for n in names])
    g["param_names"] = np.array(names)
    r.set_d2("xyz_temp", r.d2("xyz", N, 3))
    r.set_d2("F_temp", r.d2("F", N, nn))
    r.set_d1("Pex_temp", r.d1("Pex", dim * N))
    L.calcStiffness3DFiniteDifference(6)
    d_arr = (DispBCPara * 1)(DispBCPara(1, b"z", 0.0))
    f_arr = (ForceBCPara * 1)(ForceBCPara(2, b"x", 0.0, b"y", 0.0, b"z", -2000.0))
    L.setDispBC(1, d_arr)
    L.setForceBC(1, f_arr)
    # the predictor would be a no-op right after a pure force step (xyz == xyz_temp): move the top layer a little
    x = r.get("xyz")
    x[x[:, 2] > 2.4, 2] -= 1.0e-3
    x[:, 0] += 2.0e-4 * np.sin(3.0 * x[:, 1])
    r.put("xyz", x)
    phase(r, g, "s1.pred", 4, L.computeBondForceIncrementalUpdating)
    L.computeBondForceGeneral(4, 1)
    L.updateRR()
    nr, nf = r.norms()
    tol, ni = max(nr, nf), 0
    while nr > 1e-4 * tol and ni < 100:
        L.switchStateV(0)
        L.setDispBC_stiffnessUpdate3D()
        L.solverCG()
        if ni == 0:
            phase(r, g, "s1.j2", 0, L.computeBondForceJ2mixedLinear3D)
            phase(r, g, "s1.el", 6, L.computeBondForceElastic)
        L.computeBondForceGeneral(0, 1)
        L.updateRR()
        nr = r.norms()[0]
        ni += 1
    L.updateDamageGeneral(b"/dev/null", 1, 0)
    L.updateCrack()
    L.switchStateV(1)
    # load step 2, first Newton iteration: the same laws with plastic history in slot [0]
    r.set_d2("xyz_temp", r.d2("xyz", N, 3))
    r.set_d2("F_temp", r.d2("F", N, nn))
    r.set_d1("Pex_temp", r.d1("Pex", dim * N))
    L.calcStiffness3DFiniteDifference(6)
    L.setDispBC(1, d_arr)
    L.setForceBC(1, f_arr)
    L.computeBondForceGeneral(4, 1)
    L.updateRR()
    L.switchStateV(0)
    L.setDispBC_stiffnessUpdate3D()
    L.solverCG()
    phase(r, g, "s2.j2", 0, L.computeBondForceJ2mixedLinear3D)
    phase(r, g, "s2.el", 6, L.computeBondForceElastic)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_particle.npz"))
    np.savez_compressed(out, **g)
    dl = [float(g[f"s2.j2.c{k}.J2_dlambda"][PARTICLES[k]]) for k in range(len(PARTICLES))]
    print("J2_dlambda of the called particles (step 2):", dl, "| dLp[0] max before it:", np.abs(g["s2.j2.pre.dLp"][..., 0]).max())
    print("predictor |ddL| max", np.abs(g["s1.pred.c4.ddL"]).max(), "| elastic F max", np.abs(g["s2.el.c4.F"]).max(), "| Newton iterations in step 1:", ni)
    print("wrote", out, out.stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    main()
