"""GPU parity for the parts of the reference's API surface that no shipped driver exercises (SURVEY section 8, rows a8 /
a16 "unused variants"; constitutive.h:24,26): the two ductile-damage laws updateDamageGeneral's dispatcher keeps
commented out, through the C ABI and through the reference-named drop-in entry points.  Fixture:
tests/golden/sc6_damage_variants.npz (made by tests/golden/make_golden_damage_variants.py from oracle/_ref)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_same, get_slots, params_from_golden, put_slots, rel_err

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REFDIR = ROOT / "oracle" / "_ref"
GOLD = ROOT / "tests" / "golden"


def _regen(script, tmp_path, out_name):
    """run a golden generator with the reference's HOST code + GPU drop-in library instead of the all-CPU build"""
    import sys
    host = REFDIR / "liblpmc_b200host.so"
    if not host.exists():
        pytest.skip("oracle/_ref/liblpmc_b200host.so not built")
    out = tmp_path / out_name
    env = dict(os.environ, LPMB_REF_SO=str(host), LPMB_GOLDEN_OUT=str(out))
    r = subprocess.run([sys.executable, str(GOLD / script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _dv_ctx(lpm, g, pre):
    """context for tests/golden/sc6_damage_variants.npz: topology + the damage state recorded before a call"""
    N, nn = g["setup.neighbors"].shape
    c = lpm.Context(N, 3, 2, nn, g["setup.conn"].shape[1])
    c.set_params(**params_from_golden(g))
    c.set_field("xyz", g["setup.xyz"])
    c.set_field("xyz_initial", g["setup.xyz"])
    c.set_neighbors(g["setup.neighbors"], g["setup.nsign"])   # derives nb_initial, mirror slots, distance_initial
    for n in ("J2_dlambda", "J2_triaxiality", "damage_broken", "damage_w"):
        c.set_field(n, g[f"{pre}.{n}"])
    put_slots(c, "damage_D", g[f"{pre}.damage_D"])
    put_slots(c, "damage_local", g[f"{pre}.damage_local"])
    put_slots(c, "damage_nonlocal", g[f"{pre}.damage_nonlocal"])
    return c


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_particlewise_local_damage_bit_exact(lpm, step):
    """updateDuctileDamagePwiseLocal (constitutive.c:1529-1579; commented out in the reference's dispatcher, SURVEY row
    a16): calls 2 and 3 detach 9 and 16 particles, each losing all its bonds in both directions"""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sc6_damage_variants.npz")
    pre, post = f"pwl.{step}.pre", f"pwl.{step}.post"
    c = _dv_ctx(lpm, g, pre)
    assert_same(c.get_field("distance_initial"), g["setup.distance_initial"], "distance_initial")
    broken, pairs = c.update_damage(lpm.capi.DAMAGE_PWISE_LOCAL)
    assert broken == int(g[f"pwl.{step}.broken"][0])
    for n in ("damage_broken", "damage_w"):
        assert_same(c.get_field(n), g[f"{post}.{n}"], n)
    assert_same(get_slots(c, "damage_D", 2), g[f"{post}.damage_D"], "damage_D")
    assert_same(get_slots(c, "damage_local", 2), g[f"{post}.damage_local"], "damage_local")
    detached = np.flatnonzero((g[f"{post}.damage_local"][:, 0] == 1.0) & (g[f"{pre}.damage_local"][:, 0] != 1.0))
    assert np.array_equal(pairs[:, 0], detached) and (pairs[:, 1] == -1).all()
    c.close()


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_bondwise_nonlocal_damage(lpm, step):
    """updateDuctileDamageBwiseNonlocal (constitutive.c:1698-1753; commented out in the dispatcher): Gaussian average
    over the bond list (exp() on the device vs glibc: 1e-12 on damage values, bit-exact on which bonds break)"""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sc6_damage_variants.npz")
    pre, post = f"bwn.{step}.pre", f"bwn.{step}.post"
    c = _dv_ctx(lpm, g, pre)
    broken, pairs = c.update_damage(lpm.capi.DAMAGE_BWISE_NONLOCAL)
    assert broken == int(g[f"bwn.{step}.broken"][0])
    assert_same(c.get_field("damage_broken"), g[f"{post}.damage_broken"], "damage_broken")
    tol = 1e-12   # fp64 exp(): device libm vs glibc
    assert rel_err(get_slots(c, "damage_nonlocal", 2), g[f"{post}.damage_nonlocal"]) <= tol
    assert rel_err(get_slots(c, "damage_D", 2), g[f"{post}.damage_D"]) <= tol
    assert rel_err(c.get_field("damage_w"), g[f"{post}.damage_w"]) <= tol
    newly = (g[f"{pre}.damage_broken"] != 0) & (g[f"{post}.damage_broken"] == 0)
    ii, jj = np.nonzero(newly)
    assert np.array_equal(pairs, np.stack([ii, g["setup.neighbors"][ii, jj]], axis=1))   # i ascending, then slot ascending
    c.close()


def test_dropin_replays_damage_variant_golden_case(tmp_path):
    """tests/golden/sc6_damage_variants.npz: updateDuctileDamagePwiseLocal / updateDuctileDamageBwiseNonlocal (reference
    names, constitutive.h:24,26) through liblpmc_dropin.so, all three calls of each law.  The generator pokes the
    multiplier / triaxiality fields into host memory between calls and announces it with
    lpmc_dropin_invalidate_state() (include/lpmc_dropin.h)."""
    new = _regen("make_golden_damage_variants.py", tmp_path, "dv.npz")
    old = np.load(GOLD / "sc6_damage_variants.npz")
    for s in ("s1", "s2", "s3"):
        assert int(new[f"pwl.{s}.broken"][0]) == int(old[f"pwl.{s}.broken"][0])
        assert int(new[f"bwn.{s}.broken"][0]) == int(old[f"bwn.{s}.broken"][0])
        for n in ("damage_local", "damage_broken", "damage_w", "damage_D"):
            assert np.array_equal(new[f"pwl.{s}.post.{n}"], old[f"pwl.{s}.post.{n}"]), (s, n)
        assert np.array_equal(new[f"bwn.{s}.post.damage_broken"], old[f"bwn.{s}.post.damage_broken"])
        for n in ("damage_nonlocal", "damage_w", "damage_D"):
            assert _rel(new[f"bwn.{s}.post.{n}"], old[f"bwn.{s}.post.{n}"]) <= 1e-12, (s, n)


# ---- per-particle law entry points (constitutive.h:15,17,20) -------------------------------------------------------
PP_WRITES = {
    4: ("ddL", "ddL_total", "TddL_total", "F", "Pin"),
    6: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin"),
    0: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "stress_tensor", "J2_dlambda", "dLp2",
        "J2_beta2", "J2_alpha2"),
}
PP_STATE = ("dL", "dL_ave", "ddLp", "ddL", "csx", "csy", "csz", "F", "F_temp", "damage_broken", "damage_w", "dL_total", "TdL_total", "ddL_total",
            "TddL_total", "stress_tensor", "J2_dlambda", "xyz", "xyz_temp", "Pin", "pl_flag", "nb")


def _pp_ctx(lpm, g, pre):
    from helpers import make_ctx
    c = make_ctx(lpm, g)
    for n in PP_STATE:
        c.set_field(n, g[f"{pre}.{n}"])
    put_slots(c, "dLp", g[f"{pre}.dLp"])
    put_slots(c, "J2_beta", g[f"{pre}.J2_beta"])
    put_slots(c, "J2_alpha", g[f"{pre}.J2_alpha"])
    return c


@pytest.mark.parametrize("tag,law", [("s1.pred", 4), ("s1.j2", 0), ("s1.el", 6), ("s2.j2", 0), ("s2.el", 6)])
def test_per_particle_laws_bit_exact(lpm, tag, law):
    """computeBondForceIncrementalUpdating(ii) / computeBondForceJ2mixedLinear3D(ii) / computeBondForceElastic(ii) called
    outside the dispatcher, five particles in sequence (one with a broken bond, its partner, a corner, an interior one,
    one of the loaded layer): after every call EVERY array the law may write equals the reference's bit for bit -- the
    star's rows changed, all other rows did not (tests/golden/sc6_particle.npz; step 2 carries plastic history)."""
    g = np.load(GOLD / "sc6_particle.npz")
    c = _pp_ctx(lpm, g, f"{tag}.pre")
    changed = 0
    for k, ii in enumerate(g[f"{tag}.particles"]):
        c.bond_force_particle(law, int(ii))
        for n in PP_WRITES[law]:
            want = g[f"{tag}.c{k}.{n}"]
            assert_same(c.get_field(n), want, f"{tag} call {k} (particle {ii}): {n}")
            base = n[:-1] if n.endswith("2") and n[:-1] in ("dLp", "J2_beta", "J2_alpha") else n
            before = g[f"{tag}.pre.{base}"][..., 2] if base != n else g[f"{tag}.pre.{n}"]
            changed += int((np.asarray(want) != np.asarray(before)).any())
    assert changed > 0                                     # the calls really did something
    if law == 0:
        assert int(g[f"{tag}.c2.pl_flag"].sum()) > 0       # and the J2 case is plastic
    c.close()


def test_per_particle_law_rejects_other_plmodes(lpm):
    g = np.load(GOLD / "sc6_particle.npz")
    c = _pp_ctx(lpm, g, "s1.j2.pre")
    for plmode in (2, 7):
        with pytest.raises(lpm.LPMBError):
            c.bond_force_particle(plmode, 0)
    with pytest.raises(lpm.LPMBError):
        c.bond_force_particle(6, 216)
    c.close()


def test_dropin_replays_per_particle_golden_case(tmp_path):
    """tests/golden/sc6_particle.npz with every call -- including the per-particle ones, by their reference names --
    going through liblpmc_dropin.so.  Step 1 phases are reached through bit-exact calls only (FD tangent, predictor) or
    one CG solve (1e-10), so: predictor phase bit-exact, J2 / elastic phases of step 1 to 1e-9."""
    new = _regen("make_golden_particle.py", tmp_path, "pp.npz")
    old = np.load(GOLD / "sc6_particle.npz")
    for k in range(5):
        for n in PP_WRITES[4]:
            assert np.array_equal(new[f"s1.pred.c{k}.{n}"], old[f"s1.pred.c{k}.{n}"]), (k, n)
        for tag, law in (("s1.j2", 0), ("s1.el", 6)):
            for n in PP_WRITES[law]:
                assert _rel(new[f"{tag}.c{k}.{n}"], old[f"{tag}.c{k}.{n}"]) <= 1e-9, (tag, k, n)
    assert np.array_equal(new["s1.j2.c4.pl_flag"], old["s1.j2.c4.pl_flag"])


# ---- binary snapshots (checkpoint / resume) -------------------------------------------
def test_snapshot_resume_is_bit_identical(lpm, tmp_path):
    """one plastic load step, lpmb_snapshot_save, a second load step -> A; a FRESH context, lpmb_snapshot_load, the
    same second load step -> B; A == B bit for bit (state, parameters, topology, BC indices all travel in the file;
    the block pattern of K and the DoF mask are rebuilt, the tangent is re-assembled by the step itself)"""
    from helpers import make_ctx
    g = np.load(GOLD / "sc6_j2.npz")
    dbp, fbp = [(1, "z", 0.0)], [(2, 0.0, 0.0, -2000.0)]
    names = ("xyz", "F", "Pin", "stress_tensor", "damage_w", "damage_nonlocal0", "dLp0", "dLp1", "J2_alpha1", "J2_beta1", "dL", "residual", "Pex")
    c = make_ctx(lpm, g)
    c.compute_dl()
    lpm.driver.load_step(c, 0, dbp, fbp)
    snap = tmp_path / "step1.lpmb"
    c.snapshot_save(snap)
    assert snap.stat().st_size > 216 * 18 * 8 * 10
    log_a = lpm.driver.load_step(c, 0, dbp, fbp)
    A = {n: c.get_field(n) for n in names}
    N, nn, nconn = c.N, c.nn, c.nconn
    c.close()
    c2 = lpm.Context(N, 3, 2, nn, nconn)
    c2.snapshot_load(snap)
    log_b = lpm.driver.load_step(c2, 0, dbp, fbp)
    assert log_b.newton_iterations == log_a.newton_iterations == int(g["newton_counts"][1])
    assert log_b.cg_iterations == log_a.cg_iterations
    for n in names:
        assert_same(c2.get_field(n), A[n], n)
    # a context of another shape refuses the file
    c3 = lpm.Context(N + 32, 3, 2, nn, nconn)
    with pytest.raises(lpm.LPMBError):
        c3.snapshot_load(snap)
    c3.close()
    c2.close()


# ---- examples/sc_block.c: the default problem device-resident from plain C ----------------------------------------
def test_c_example_reproduces_default_case_known_answers(tmp_path):
    """examples/sc_block.c (C host code over the C ABI only, O(N) device set-up) on n = 21 = the default case C1: Newton
    iterations 2 2 1, CG iterations 80 / 106 in load step 1 and the step-1 mean displacement of the loaded layer
    (SURVEY section 8(c) known answers; tests/golden/c1_result_disp.txt line 2); binary snapshot written."""
    import re
    exe = ROOT / "examples" / "sc_block"
    if not exe.exists():
        pytest.skip("examples/sc_block not built (run __graft_entry__.build())")
    r = subprocess.run([str(exe), "21", "3", "3", str(tmp_path / "c1")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Particle number is 9261, stiffness matrix size is 2203713" in r.stdout
    steps = re.findall(r"Loading step (\d+) has finished in (\d+) iterations; CG iterations:([ \d]+);", r.stdout)
    assert [int(s[1]) for s in steps] == [2, 2, 1], r.stdout
    cg1 = [int(x) for x in steps[0][2].split()]
    # the lattice here starts at (-0.2, -0.2, -0.2); the reference's createCuboid shifts y and z by its box padding
    # (initialization.c:240-284): same problem up to a translation, so CG may stop an iteration earlier or later
    assert abs(cg1[0] - 80) <= 1 and abs(cg1[1] - 106) <= 1, cg1
    assert (tmp_path / "c1_step0003.lpmb").stat().st_size > 9261 * 18 * 8 * 10
    r1 = subprocess.run([str(exe), "21", "1"], capture_output=True, text=True, timeout=300)
    uz = float(re.search(r"mean z-displacement of the loaded layer after 1 steps: (\S+)", r1.stdout).group(1))
    gold = float((GOLD / "c1_result_disp.txt").read_text().split("\n")[1].split()[1])
    assert gold == pytest.approx(-1.27857453e-03, rel=1e-8)
    assert uz == pytest.approx(gold, rel=1e-6)


# ---- the 2-D configurations pinned by committed golden vectors (hexagonal / square lattice, brittle) ----------------
def _ctx_2d(lpm, g):
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    c = lpm.Context(N, 2, int(par["lattice"]), nn, g["setup.conn"].shape[1])
    c.set_params(radius=par["radius"], particle_volume=par["particle_volume"], critical_bstrain=par["critical_bstrain"], nbreak=par["nbreak"])
    c.set_field("xyz", g["setup.xyz"])
    c.set_field("xyz_initial", g["setup.xyz"])
    c.set_neighbors(g["setup.neighbors"], g["setup.nsign"])
    c.set_connectivity(g["setup.conn"])
    c.set_field("type", g["setup.type"])
    c.calc_kntv(g["setup.Ce"])
    return c


@pytest.mark.parametrize("name", ["hex2d_brittle", "sq2d_brittle"])
def test_2d_setup_tangent_and_laws_bit_exact(lpm, name):
    """tests/golden/{hex,sq}2d_brittle.npz (BASELINE configs 2 and 3 on a small box): calcKnTv of the 2-D lattices, the
    derived geometry, computedL, the 2-D FD tangent in the reference's CSR layout (first assembly and the one after
    bonds broke), predictor and elastic law -- bit for bit"""
    g = np.load(GOLD / f"{name}.npz")
    c = _ctx_2d(lpm, g)
    assert_same(c.get_field("Kn"), g["setup.Kn"], "Kn"); assert_same(c.get_field("Tv"), g["setup.Tv"], "Tv")
    for n in ("distance_initial", "csx_initial", "csy_initial"):
        assert_same(c.get_field(n), g[f"setup.{n}"], n)
    c.compute_dl()
    for n in ("dL", "dL_total", "TdL_total", "csx", "csy"):
        assert_same(c.get_field(n), g[f"setup.{n}"], n)
    for s, broken in (("s1", None), ("s4", g["s3.end.damage_broken"])):
        c.set_field("xyz", g[f"{s}.pre.xyz"])
        if broken is not None:
            c.set_field("damage_broken", broken)
        c.fd_stiffness(True)
        K, IK, JK = c.matrix_to_upper_csr()
        assert np.array_equal(IK, g[f"{s}.fd.IK"]) and np.array_equal(JK, g[f"{s}.fd.JK"])
        assert_same(K, g[f"{s}.fd.K_global"], f"{s} K_global")
        for n in ("dL", "csx", "csy", "dL_total", "TdL_total", "F"):
            assert_same(c.get_field(n), g[f"{s}.fd.{n}"], f"{s} FD side effect {n}")
    # predictor + elastic law on the recorded states of step 1
    c.set_field("damage_broken", np.ones_like(g["setup.Kn"]))
    for n in ("dL", "csx", "csy", "csz", "F", "dL_total", "TdL_total"):
        c.set_field(n, g[f"s1.fd.{n}"])
    c.set_field("xyz", g["s1.bc.xyz"])
    c.set_field("xyz_temp", g["s1.pre.xyz"])
    c.set_field("F_temp", g["s1.pre.F"])
    c.bond_force(4)
    for n in ("ddL", "ddL_total", "TddL_total", "F", "Pin"):
        assert_same(c.get_field(n), g[f"s1.pred.{n}"], f"predictor {n}")
    c.set_field("xyz", g["s1.e0.n0.xyz"])
    c.bond_force(6)
    for n in ("dL", "csx", "csy", "dL_total", "TdL_total", "F", "Pin", "stress_tensor", "bond_stress"):
        assert_same(c.get_field(n), g[f"s1.e0.n0.bf.{n}"], f"elastic law {n}")
    c.close()


@pytest.mark.parametrize("name,steps", [("hex2d_brittle", 3), ("sq2d_brittle", 4)])
def test_2d_brittle_trajectory(lpm, name, steps):
    """the same cases as whole load steps, device-resident (driver.py over the C ABI): Newton iterations, the number of
    breaking events and WHICH bonds break (candidates > nbreak -> the reference's shell-sort selection) identical,
    displacements and bond forces to 1e-9"""
    g = np.load(GOLD / f"{name}.npz")
    c = _ctx_2d(lpm, g)
    c.compute_dl()
    dbp = [(1, "x", 1.5e-4), (1, "y", 0.0), (2, "x", 0.0), (2, "y", 0.0)]
    x0 = g["setup.xyz"]
    for step in range(1, steps + 1):
        log = lpm.driver.load_step(c, 6, dbp, [])
        s = f"s{step}"
        assert log.newton_iterations == int(g["newton_counts"][step - 1]), (step, log.newton_iterations)
        assert log.reassemblies + 1 == int(g[f"{s}.events"][0])
        assert abs(log.cg_iterations[0] - int(g[f"{s}.e0.n0.cg_iters"][0])) <= 1
        assert_same(c.get_field("damage_broken"), g[f"{s}.end.damage_broken"], f"{s} damage_broken")
        assert rel_err(c.get_field("xyz") - x0, g[f"{s}.end.xyz"] - x0) <= 1e-9
        assert rel_err(c.get_field("F"), g[f"{s}.end.F"]) <= 1e-9
    assert (g[f"s{steps}.end.damage_broken"] == 0).sum() > 0
    c.close()


# ---- BCC lattice (8 + 6 neighbours, 41 conn, 24 slip systems): crystal plasticity on the reference's fifth lattice --------
def test_bcc_crystal_plasticity():
    """runs impl_bcc_crystal_plasticity in a child process (child process: keeps a failure of the library there)"""
    from helpers import run_isolated
    print(run_isolated(__file__, "impl_bcc_crystal_plasticity")[-300:])


def impl_bcc_crystal_plasticity(lpm):
    """tests/golden/bcc_cp.npz (tests/golden/make_golden_cp.py with LPMB_CP_LATTICE=4): topology from the O(N) device
    builder, calcKnTv of the BCC lattice (stiffness.c:236-266), computeCab bit-exact; two load steps of the Miehe law
    device-resident: Newton iteration counts and active slip systems identical, displacements 1e-9"""
    g = np.load(GOLD / "bcc_cp.npz")
    par = {str(k): float(v) for k, v in zip(g["param_names"], g["params"])}
    N, nn = g["setup.neighbors"].shape
    assert (nn, g["setup.conn"].shape[1], int(par["nslipSys"])) == (14, 41, 24)
    c = lpm.Context(N, 3, 4, nn, g["setup.conn"].shape[1])
    c.set_params(**{k: v for k, v in par.items() if k != "nslipSys"})
    c.set_field("xyz", g["setup.xyz"])
    c.set_field("xyz_initial", g["setup.xyz"])
    c.build_topology(par["neighbor1_cutoff"], par["neighbor2_cutoff"])
    c.set_field("type", g["setup.type"])
    c.calc_kntv(g["setup.Ce"])
    c.compute_dl()
    c.set_schmid_tensor(g["setup.schmid_tensor"])
    put_slots(c, "cp_gy", g["setup.cp_gy"])
    assert_same(c.get_field("neighbors"), g["setup.neighbors"], "neighbors")
    assert_same(c.get_field("nsign"), g["setup.nsign"], "nsign")
    assert np.array_equal(c.k_pointer(), g["setup.K_pointer"])
    assert_same(c.get_field("Kn"), g["setup.Kn"], "Kn")
    assert_same(c.get_field("Tv"), g["setup.Tv"], "Tv")
    c.compute_cab()
    assert_same(c.get_field("cp_Cab"), g["setup.cp_Cab"], "cp_Cab")
    dbp = [(int(t), str(a), float(v)) for t, a, v in zip(g["dbp_type"], g["dbp_axis"], g["dbp_step"])]
    for step in (1, 2):
        log = lpm.driver.load_step(c, 1, dbp, [])
        assert log.newton_iterations == int(g["newton_counts"][step - 1])
        s = f"s{step}.end"
        u, u_ref = c.get_field("xyz") - g["setup.xyz"], g[f"{s}.xyz"] - g["setup.xyz"]
        assert rel_err(u, u_ref) <= 1e-9
        assert rel_err(c.get_field("F"), g[f"{s}.F"]) <= 1e-8
        assert np.array_equal(c.get_field("cp_Jact"), g[f"{s}.cp_Jact"])
    c.close()
