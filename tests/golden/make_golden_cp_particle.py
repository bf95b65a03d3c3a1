"""Generate tests/golden/fcc_cp_particle.npz from the UNMODIFIED reference (oracle/_ref), single thread.

computeBondForceCPMiehe(ii) (src/constitutive.c:866-1396, include/constitutive.h:19) called directly, outside the
computeBondForceGeneral dispatcher.  Its result depends on the memo state_v (constitutive.c:946-959): a star member whose
flag is 0 is return-mapped and flagged, one whose flag is 1 REUSES the increments (ddLp, cp_dgy, cp_dA, cp_dA_single) an
earlier call left.  The dispatcher zeroes the memo before its serial loop (:116) and leaves it all 1.

Case: the 256-particle FCC block of fcc_cp.npz (material / BCs of examples/FCC_Al_R0.3_001_tension.c).
  fresh   after the first CG solve of load step 1, memo zeroed by hand (what the dispatcher's memset does): five calls in
          sequence -- a particle, one of its neighbours (overlapping stars: part of the second star is reused), a far
          particle, a corner, the first particle again (everything reused)
  memo    after computeBondForceGeneral(1, .) (memo all 1) + updateRR + the next CG solve: three calls that reuse the
          previous iteration's increments on the NEW positions
For every phase: the state before the first call and, after each call, every array the law may write.
Run here (container with /root/reference):   python tests/golden/make_golden_cp_particle.py
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import RefLPM  # noqa: E402

DBP = [(1, "z", 0.0), (2, "x", 0.0), (2, "z", 0.0), (3, "y", 0.0), (3, "z", 0.0), (4, "x", 0.0), (4, "y", 0.0), (4, "z", 0.0),
       (5, "z", -2.0e-3)]
BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "xyz", "Pin", "pl_flag", "nb", "state_v")
CP = ("cp_gy", "cp_A", "cp_A_single", "cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single")
WRITES = ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "state_v")
WRITES_CP = ("cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single")


def state(r, prefix, g):
    for n in BOND + PART:
        g[f"{prefix}.{n}"] = r.get(n)
    g[f"{prefix}.dLp"] = r.get("dLp")
    for n in CP:
        g[f"{prefix}.{n}"] = r.get_cp(n)


def phase(r, g, tag, particles):
    if hasattr(r.lib, "lpmc_dropin_invalidate_state"):   # replay through the GPU drop-in layer
        r.lib.lpmc_dropin_invalidate_state()
    state(r, f"{tag}.pre", g)
    for k, ii in enumerate(particles):
        r.lib.computeBondForceCPMiehe(int(ii))
        for n in WRITES:
            g[f"{tag}.c{k}.{n}"] = r.get(n)
        for n in WRITES_CP:
            g[f"{tag}.c{k}.{n}"] = r.get_cp(n)
        g[f"{tag}.c{k}.dLp2"] = r.get("dLp")[..., 2]
        g[f"{tag}.c{k}.cp_gy2"] = r.get_cp("cp_gy")[..., 2]
        g[f"{tag}.c{k}.cp_A_single2"] = r.get_cp("cp_A_single")[..., 2]
        g[f"{tag}.c{k}.cp_A2"] = r.get_cp("cp_A")[:, 2]
    g[f"{tag}.particles"] = np.array(particles)


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_fcc(box=(0, 3.5, 0, 3.5, 0, 3.5), lattice=3)
    L = r.lib
    N = r.N
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
    nbr = r.get("neighbors")
    xyz = r.get("xyz")
    centre = int(np.argmin(((xyz - xyz.mean(axis=0)) ** 2).sum(axis=1)))
    neighbour = int(nbr[centre, 0])
    far = int(np.argmax(((xyz - xyz[centre]) ** 2).sum(axis=1)))
    fresh = [centre, neighbour, far, 0, centre]
    # step 1 runs as the driver would (plastic from the first iteration on this case); step 2 carries the history
    for step in (1, 2):
        nr, nf = r.begin_step(DBP, [])
        tol, ni = max(nr, nf), 0
        while nr > 1e-4 * tol and ni < 100:
            L.switchStateV(0)
            L.setDispBC_stiffnessUpdate3D()
            L.solverCG()
            if step == 2 and ni == 0:
                r.put("state_v", np.zeros(N, dtype=np.int32))      # the dispatcher's memset, constitutive.c:116
                phase(r, g, "fresh", fresh)
            if step == 2 and ni == 1:
                assert int(r.get("state_v").min()) == 1            # left by the dispatcher pass of iteration 0
                phase(r, g, "memo", [centre, far, 0])
            L.computeBondForceGeneral(1, 1)
            L.updateRR()
            nr = r.norms()[0]
            ni += 1
        assert step == 1 or ni >= 2, ni
        L.computeStrain()
        L.updateDamageGeneral(b"/dev/null", step, 1)
        L.updateCrack()
        L.switchStateV(1)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "fcc_cp_particle.npz"))
    np.savez_compressed(out, **g)
    for tag in ("fresh", "memo"):
        ps = g[f"{tag}.particles"]
        act = [int(g[f"{tag}.c{k}.cp_Jact"][ps[k]].sum()) for k in range(len(ps))]
        flagged = [int(g[f"{tag}.c{k}.state_v"].sum()) for k in range(len(ps))]
        print(tag, "particles", ps.tolist(), "active slip systems of the called particle", act, "flagged particles after each call", flagged,
              "finite", all(np.isfinite(g[f"{tag}.c{k}.F"]).all() for k in range(len(ps))))
    print("wrote", out, round(out.stat().st_size / 1e6, 2), "MB")


if __name__ == "__main__":
    main()
