"""Generate tests/golden/sc6_particle2.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so), single thread.

The two remaining J2 PER-PARTICLE law entry points (include/constitutive.h:18,21) called directly, outside the
computeBondForceGeneral dispatcher:
    computeBondForceJ2energyReturnMap(ii, t)   src/constitutive.c:286-463   (plmode 3)
    computeBondForceJ2nonlinearIso(ii)         src/constitutive.c:689-863   (plmode 5)
One call rewrites the geometry / return-map outputs of ii AND of its neighbours across intact bonds and the force
outputs of ii; plmode 5 also advances the plastic state of the whole star IN PLACE (slot [0]) and leaves the trial bond
forces of the star members in F.  Across a broken bond the force pass of ii reads whatever the partner's rows hold
(the partner is not in the star, so it is not refreshed).

Cases (the 6^3 blocks of sc6_j2energy.npz / sc6_j2iso.npz, same material and loading):
  e.s1   plmode 3, after the first CG solve of load step 1 (plastic), t = +1
  e.s2   plmode 3, load step 2 after the state was committed, t = -1, three bonds broken + updateCrack (nb < nb_initial)
  i.s1   plmode 5, after the first CG solve of load step 1 (plastic)
  i.s2   plmode 5, load step 2 with plastic history in slot [0] and the same three bonds broken
For every phase: the complete state before the first call and, after each of the five calls (made in sequence, so later
calls start from what the earlier ones left), every array the law may write.
Run here (container with /root/reference):   python tests/golden/make_golden_particle2.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import DispBCPara, ForceBCPara, RefLPM  # noqa: E402

BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "J2_dlambda", "J2_alpha", "J2_beta_eq", "xyz", "Pin", "pl_flag", "nb")
WRITES = {
    3: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "J2_dlambda", "dLp2", "J2_beta_eq2",
        "J2_alpha2"),
    5: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "dL_ave", "F", "Pin", "stress_tensor", "J2_dlambda", "dLp0", "J2_beta0",
        "J2_alpha0"),
}
BROKEN = ((100, 0), (57, 5), (130, 11))      # (particle, slot), broken in both directions as in make_golden_j2e.py
# 100 and 57 own a broken bond; 94 is the partner across 100's (slot 0 = -x... whichever the list gives); corner; interior
PARTICLES = [100, 0, 129, 215, 57]


def state(r, prefix, out):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.dLp"] = r.get("dLp")
    out[f"{prefix}.J2_beta"] = r.get("J2_beta")
    out[f"{prefix}.damage_D"] = r.get("damage_D")


def written(r, n):
    if n[-1] in "02" and n[:-1] in ("dLp", "J2_beta", "J2_alpha", "J2_beta_eq"):
        return r.get(n[:-1])[..., int(n[-1])]
    return r.get(n)


def phase(r, g, tag, law, fn, particles):
    if hasattr(r.lib, "lpmc_dropin_invalidate_state"):   # replay through the GPU drop-in layer
        r.lib.lpmc_dropin_invalidate_state()
    state(r, f"{tag}.pre", g)
    for k, ii in enumerate(particles):
        fn(ii)
        for n in WRITES[law]:
            g[f"{tag}.c{k}.{n}"] = written(r, n)
    g[f"{tag}.particles"] = np.array(particles)


def break_bonds(r):
    nbr, br, w = r.get("neighbors"), r.get("damage_broken"), r.get("damage_w")
    partners = []
    for i, j in BROKEN:
        k = int(nbr[i, j])
        jj = int(np.where(nbr[k] == i)[0][0])
        br[i, j] = br[k, jj] = 0.0
        w[i, j] = w[k, jj] = 0.0
        partners.append(k)
    r.put("damage_broken", br)
    r.put("damage_w", w)
    if hasattr(r.lib, "lpmc_dropin_invalidate_state"):   # drop-in replay: a HOST edit of device-authoritative state must be announced
        r.lib.lpmc_dropin_invalidate_state()             # (include/lpmc_dropin.h, "State ownership"); without it the device keeps nb = 18
    r.lib.updateCrack()
    return partners


def run(g, law, tagp, setup_kw, extra, loads, clamp_all):
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5, plmode=law, **setup_kw)
    for k, v in extra.items():
        r.sd(k, v)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    assert N == 216, N
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "type", "distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn",
              "Tv", "Ce", "sigmay"):
        g[f"{tagp}.setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "J2_H", "J2_xi", "J2_C", "damage_L", "damage_threshold", "damagec_A", "neighbor1_cutoff",
             "neighbor2_cutoff"]
    g[f"{tagp}.params"] = np.array([r.gd(n) for n in names])
    g[f"{tagp}.param_names"] = np.array(names)
    particles = list(PARTICLES)
    for step, (t, load) in enumerate(loads, start=1):
        if step == 2:
            partners = break_bonds(r)
            particles = [PARTICLES[0], partners[0], PARTICLES[2], PARTICLES[3], PARTICLES[4]]   # 100, its partner, ..., 57
        r.set_d2("xyz_temp", r.d2("xyz", N, 3))
        r.set_d2("F_temp", r.d2("F", N, nn))
        r.set_d1("Pex_temp", r.d1("Pex", dim * N))
        L.calcStiffness3DFiniteDifference(6)
        if clamp_all:
            d_arr = (DispBCPara * 3)(DispBCPara(1, b"x", 0.0), DispBCPara(1, b"y", 0.0), DispBCPara(1, b"z", 0.0))
            L.setDispBC(3, d_arr)
        else:
            d_arr = (DispBCPara * 1)(DispBCPara(1, b"z", 0.0))
            L.setDispBC(1, d_arr)
        f_arr = (ForceBCPara * 1)(ForceBCPara(2, b"x", 0.0, b"y", 0.0, b"z", load))
        L.setForceBC(1, f_arr)
        L.computeBondForceGeneral(4, t)
        L.updateRR()
        nr, nf = r.norms()
        tol, ni = max(nr, nf), 0
        while nr > 1e-4 * tol and ni < 2:
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            if ni == 0:
                if law == 3:
                    phase(r, g, f"{tagp}.s{step}", 3, lambda ii: L.computeBondForceJ2energyReturnMap(ii, t), particles)
                else:
                    phase(r, g, f"{tagp}.s{step}", 5, L.computeBondForceJ2nonlinearIso, particles)
                g[f"{tagp}.s{step}.t"] = np.array([t])
            L.computeBondForceGeneral(law, t)
            L.updateRR()
            nr = r.norms()[0]
            ni += 1
        L.switchStateV(1)


def main():
    g: dict = {}
    run(g, 3, "e", dict(J2_xi=0.3), {}, ((1, -2000.0), (-1, 2000.0)), clamp_all=False)
    run(g, 5, "i", dict(damagec_A=1.5, damage_threshold=0.02), {"J2_C": 2000.0}, ((1, -9000.0), (1, -3000.0)), clamp_all=True)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_particle2.npz"))
    np.savez_compressed(out, **g)
    for tag in ("e.s1", "e.s2", "i.s1", "i.s2"):
        ps = g[f"{tag}.particles"]
        dl = [float(g[f"{tag}.c{k}.J2_dlambda"][ps[k]]) for k in range(len(ps))]
        fin = all(np.isfinite(g[f"{tag}.c{k}.F"]).all() for k in range(len(ps)))
        print(tag, "particles", ps.tolist(), "dlambda of the called particles", dl, "finite F", fin, "nb min", int(g[f"{tag}.pre.nb"].min()))
    print("wrote", out, out.stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    main()
