"""Generate tests/golden/sc6_damage_variants.npz from the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so), single thread.

The two ductile-damage laws that updateDamageGeneral's dispatcher keeps commented out (src/constitutive.c:155-157):
  updateDuctileDamagePwiseLocal     (:1529-1604)  particle-wise local accumulation; a particle beyond the threshold is
                                                   set to 1 and loses ALL its bonds (both directions)
  updateDuctileDamageBwiseNonlocal  (:1698-1753)  Gaussian average over the bond list (DAM_PHI(distance_initial), self
                                                   weight = particle_volume with phi omitted), bond breaks when the MEAN of
                                                   its two end values passes the threshold
They only read J2_dlambda / J2_triaxiality and the damage state, so the fixture feeds synthetic multiplier and
triaxiality fields on the 6^3 block of sc6_j2.npz (one bond pre-broken) and records every call's inputs and outputs;
three calls per law so that the frozen / clamped branches and already-broken bonds are exercised.
Run here (container with /root/reference):   python tests/golden/make_golden_damage_variants.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref import RefLPM  # noqa: E402

FIELDS = ("J2_dlambda", "J2_triaxiality", "damage_local", "damage_nonlocal", "damage_broken", "damage_w", "damage_D")


def state(r, prefix, out):
    for n in FIELDS:
        out[f"{prefix}.{n}"] = r.get(n)


def fields(x0, step):
    """synthetic plastic-multiplier / triaxiality fields: a hot spot that moves with the step, and a region of
    strongly negative triaxiality where the (1 + A triax) factor switches the accumulation off"""
    c = np.array([1.25, 1.25, 1.25]) + 0.25 * (step - 1) * np.array([1.0, -1.0, 0.5])
    dl = 0.012 * step * np.exp(-1.5 * ((x0 - c) ** 2).sum(axis=1))
    triax = 0.4 * np.sin(1.7 * x0[:, 0] + 0.3 * step) * np.cos(0.9 * x0[:, 1]) - 0.5 * (x0[:, 2] < 0.6)
    return dl, triax


def run_law(r, g, tag, fn, x0):
    N, nn = r.N, r.nn
    # fresh damage state, one bond (both directions) already broken
    r.put("damage_broken", np.ones((N, nn)))
    r.put("damage_w", np.ones((N, nn)))
    r.put("damage_D", np.zeros((N, nn, 2)))
    r.put("damage_local", np.zeros((N, 2)))
    r.put("damage_nonlocal", np.zeros((N, 2)))
    nbr = r.get("neighbors")
    i0 = 86
    j0 = int(nbr[i0, 3])
    b = np.ones((N, nn))
    b[i0, 3] = 0.0
    b[j0, list(nbr[j0]).index(i0)] = 0.0
    r.put("damage_broken", b)
    ks = []
    for step in (1, 2, 3):
        dl, triax = fields(x0, step)
        r.put("J2_dlambda", dl)
        r.put("J2_triaxiality", triax)
        state(r, f"{tag}.s{step}.pre", g)
        if hasattr(r.lib, "lpmc_dropin_invalidate_state"):   # replay through the GPU drop-in layer (tests/test_dropin_gpu.py)
            r.lib.lpmc_dropin_invalidate_state()
        k = fn(b"/dev/null", step)
        g[f"{tag}.s{step}.broken"] = np.array([k])
        state(r, f"{tag}.s{step}.post", g)
        ks.append(int(k))
    return ks


def main():
    r = RefLPM.instance()
    r.threads(1)
    r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5, plmode=0, damagec_A=1.5, damage_threshold=0.02, damage_L=0.6)
    L = r.lib
    N, nn = r.N, r.nn
    assert N == 216, N
    g: dict = {}
    for n in ("xyz", "neighbors", "nsign", "nb_initial", "conn", "distance_initial"):
        g[f"setup.{n}"] = r.get(n)
    names = ["radius", "particle_volume", "damage_L", "damage_threshold", "damagec_A", "neighbor1_cutoff", "neighbor2_cutoff"]
    g["params"] = np.array([r.gd(n) for n in names])
    g["param_names"] = np.array(names)
    x0 = g["setup.xyz"]
    k1 = run_law(r, g, "pwl", L.updateDuctileDamagePwiseLocal, x0)
    k2 = run_law(r, g, "bwn", L.updateDuctileDamageBwiseNonlocal, x0)
    out = Path(os.environ.get("LPMB_GOLDEN_OUT", Path(__file__).resolve().parent / "sc6_damage_variants.npz"))
    np.savez_compressed(out, **g)
    print("particle-wise local: broken particles per call", k1, "damage_local max", g["pwl.s3.post.damage_local"][:, 0].max(),
          "intact bonds left", int((g["pwl.s3.post.damage_broken"] > 0).sum()))
    print("bond-wise nonlocal: broken (directed) bonds per call", k2, "damage_nonlocal max", g["bwn.s3.post.damage_nonlocal"][:, 0].max(),
          "frozen particles", int((g["bwn.s3.pre.damage_nonlocal"][:, 0] > 0.02).sum()))
    print("wrote", out, out.stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    main()
