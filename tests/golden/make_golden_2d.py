"""Generate tests/golden/hex2d_brittle.npz and sq2d_brittle.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so).

The 2-D configurations of BASELINE.json (examples/shear_hex_brittle.c: hexagonal lattice, 12 neighbours, 31 conn;
examples/3_point_bending_sq_brittle.c: square lattice, 8 neighbours, 17 conn) on a small box (110 / 100 particles): elastic
law (plmode 6) + updateBrittleDamage with nbreak = 2, sheared through the top row like the hex example (:257-266).  Per
load step: the 2-D FD tangent (calcStiffness2DFiniteDifference, stiffness.c:271-381: K_global / IK / JK with the 2x2
block layout) and its side effects, BCs, predictor, residual, every Newton iteration (BC-modified tangent, CG solve,
elastic bond force), then the damage / crack / re-assembly loop of lpmc_project.c:469-541 with every breaking event
(candidate count, which bonds the reference's shell sort selects).  critical_bstrain is set so that bonds start to break
in step 3 and more than nbreak candidates occur.
Run here (container with /root/reference):   python tests/golden/make_golden_2d.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import DispBCPara, ForceBCPara, RefLPM  # noqa: E402

BOND = ("dL", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "xyz", "Pin", "nb")


def state(r, prefix, out):
    for n in BOND + PART:
        out[f"{prefix}.{n}"] = r.get(n)
    out[f"{prefix}.damage_D"] = r.get("damage_D")


def make(r, lattice, crit, out_path):
    r.setup_2d(lattice=lattice, critical_bstrain=crit)
    L = r.lib
    N, nn, dim = r.N, r.nn, r.dim
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "nb_conn", "K_pointer", "type", "distance_initial", "csx_initial",
              "csy_initial", "csz_initial", "Kn", "Tv", "Ce", "KnTve", "distance", "dL", "dL_total", "TdL_total", "csx", "csy", "csz"):
        g[f"setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "critical_bstrain", "neighbor1_cutoff", "neighbor2_cutoff"]
    g["params"] = np.array([r.gd(n) for n in names] + [float(r.gi("nbreak")), float(lattice)])
    g["param_names"] = np.array(names + ["nbreak", "lattice"])
    d_arr = (DispBCPara * 4)(DispBCPara(1, b"x", 1.5e-4), DispBCPara(1, b"y", 0.0), DispBCPara(2, b"x", 0.0), DispBCPara(2, b"y", 0.0))
    f_arr = (ForceBCPara * 1)()
    newton_counts, events = [], []
    for step in (1, 2, 3, 4):
        s = f"s{step}"
        g[f"{s}.pre.xyz"] = r.get("xyz")
        g[f"{s}.pre.F"] = r.get("F")
        r.set_d2("xyz_temp", r.d2("xyz", N, 3))
        r.set_d2("F_temp", r.d2("F", N, nn))
        r.set_d1("Pex_temp", r.d1("Pex", dim * N))
        L.calcStiffness2DFiniteDifference(6)
        for n in ("K_global", "IK", "JK"):
            g[f"{s}.fd.{n}"] = r.get(n)
        state(r, f"{s}.fd", g)
        L.setDispBC(4, d_arr)
        L.setForceBC(0, f_arr)
        g[f"{s}.bc.xyz"] = r.get("xyz")
        g[f"{s}.bc.Pex"] = r.get("Pex")
        g[f"{s}.bc.dispBC_index"] = r.get("dispBC_index")
        g[f"{s}.bc.fix_index"] = r.get("fix_index")
        L.computeBondForceGeneral(4, 1)
        state(r, f"{s}.pred", g)
        for n in ("ddL", "ddL_total", "TddL_total"):
            g[f"{s}.pred.{n}"] = r.get(n)
        event, total_ni = 0, 0
        while True:                                        # label_broken_bond, lpmc_project.c:408
            L.updateRR()
            nr, nf = r.norms()
            if event == 0:
                g[f"{s}.rr.residual"] = r.get("residual")
                g[f"{s}.rr.norms"] = np.array([nr, nf])
            tol, ni = max(nr, nf), 0
            while nr > 1e-4 * tol and ni < 100:
                t = f"{s}.e{event}.n{ni}"
                L.switchStateV(0)
                L.setDispBC_stiffnessUpdate2D()
                rec = event == 0 and ni < 2
                if rec:
                    g[f"{t}.K_bc"] = r.get("K_global")
                    g[f"{t}.rhs"] = r.get("residual")
                L.solverCG()
                if rec:
                    g[f"{t}.disp"] = r.get("disp")
                    g[f"{t}.cg_iters"] = np.array([L.lpmb_shim_last_itercount()])
                    g[f"{t}.xyz"] = r.get("xyz")
                L.computeBondForceGeneral(6, 1)
                if rec:
                    state(r, f"{t}.bf", g)
                L.updateRR()
                nr = r.norms()[0]
                if rec:
                    g[f"{t}.residual"] = r.get("residual")
                ni += 1
            total_ni += ni
            d = f"{s}.dam{event}"
            if event == 0:
                state(r, f"{d}.pre", g)
            else:                                           # later events: only what the damage update reads
                for n in ("dL", "damage_broken", "damage_w", "F", "csx", "csy", "csz", "nb"):
                    g[f"{d}.pre.{n}"] = r.get(n)
                g[f"{d}.pre.damage_D"] = r.get("damage_D")
            broken = L.updateDamageGeneral(b"/dev/null", step, 6)
            g[f"{d}.broken"] = np.array([broken])
            for n in ("damage_broken", "damage_w", "damage_D"):
                g[f"{d}.post.{n}"] = r.get(n)
            L.updateCrack()
            for n in ("F", "Pin", "nb", "damage_visual", "fix_index"):
                g[f"{d}.crack.{n}"] = r.get(n)
            L.switchStateV(1)
            events.append((step, event, int(broken)))
            if broken <= 0:
                break
            L.calcStiffness2DFiniteDifference(6)            # :525-541
            if event == 0:
                g[f"{s}.refd.K_global"] = r.get("K_global")
            event += 1
            if event >= 6:                                  # bounded fixture: stop following the crack here
                break
        g[f"{s}.events"] = np.array([event + 1])
        g[f"{s}.end.xyz"] = r.get("xyz")
        g[f"{s}.end.F"] = r.get("F")
        g[f"{s}.end.damage_broken"] = r.get("damage_broken")
        newton_counts.append(total_ni)
    g["newton_counts"] = np.array(newton_counts)
    np.savez_compressed(out_path, **g)
    print("lattice", lattice, "N", N, "Newton iterations per step", newton_counts, "damage events (step, event, candidates)", events,
          "| intact bonds left", int(((g["s4.end.damage_broken"] > 0) & (g["setup.neighbors"] >= 0)).sum()), "of", int((g["setup.neighbors"] >= 0).sum()),
          "| wrote", out_path, round(Path(out_path).stat().st_size / 1e6, 3), "MB")


def main():
    r = RefLPM.instance()
    r.threads(1)
    here = Path(__file__).resolve().parent
    out_dir = Path(os.environ.get("LPMB_GOLDEN_OUT_DIR", here))
    make(r, 1, float(os.environ.get("LPMB_2D_CRIT_HEX", 7.0e-3)), out_dir / "hex2d_brittle.npz")
    make(r, 0, float(os.environ.get("LPMB_2D_CRIT_SQ", 7.0e-3)), out_dir / "sq2d_brittle.npz")


if __name__ == "__main__":
    main()
