"""GPU parity: bond-wise constitutive update, stress, residual, FD tangent assembly vs the reference's own
functions (golden vectors made by tests/golden/make_golden.py from oracle/_ref).

Everything here is + - * / sqrt in fp64, compiled -fmad=false, and the reference was compiled
-ffp-contract=off, so the bar is BIT-EXACT (helpers.assert_same), far inside the 1e-9 of north_star.
"""
import numpy as np
import pytest

from helpers import (assert_same, get_slots, make_ctx, put_slots, put_state, rel_err)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx(lpm, golden):
    c = make_ctx(lpm, golden)
    yield c
    c.close()


def test_calc_kntv_bit_exact(ctx, golden):
    """calcKnTv(), stiffness.c:145-204 (SC: average of the two end particles' types)"""
    assert_same(ctx.get_field("Kn"), golden["setup.Kn"], "Kn")
    assert_same(ctx.get_field("Tv"), golden["setup.Tv"], "Tv")


def test_initial_geometry_and_derived_topology(lpm, golden):
    """distance_initial / cs*_initial as searchNormalNeighbor leaves them (neighbor.c:23-26), nb_initial"""
    c = make_ctx(lpm, golden, upload_initial_geometry=False)
    for n in ("distance_initial", "csx_initial", "csy_initial", "csz_initial"):
        assert_same(c.get_field(n), golden[f"setup.{n}"], n)
    assert_same(c.get_field("nb_initial"), golden["setup.nb_initial"], "nb_initial")
    # mirror slots really are the reverse bonds
    nbr, mir = golden["setup.neighbors"], c.get_field("mirror")
    i, j = np.nonzero(nbr >= 0)
    assert np.all(nbr[nbr[i, j], mir[i, j]] == i)
    c.close()


def test_compute_dl_bit_exact(ctx, golden):
    """computedL(), lpm_basic.c:252-291"""
    ctx.compute_dl()
    for n in ("distance", "dL", "dL_total", "TdL_total", "csx", "csy", "csz"):
        assert_same(ctx.get_field(n), golden[f"setup.{n}"], n)


@pytest.mark.parametrize("step", [1, 2])
def test_fd_stiffness_bit_exact(ctx, golden, step):
    """calcStiffness3DFiniteDifference(6): K_global / IK / JK, plus the state the assembly leaves behind
    (dL, cs*, dL_total, TdL_total of the last toucher; F, Pin of the last perturbation)"""
    s = f"s{step}"
    ctx.set_field("xyz", golden[f"{s}.pre.xyz"])
    put_slots(ctx, "dLp", golden[f"{s}.pre.dLp"])
    if step == 2:
        for n in ("damage_broken", "damage_w"):
            ctx.set_field(n, golden[f"s1.crack.{n}"])
    ctx.fd_stiffness(emulate_side_effects=True)
    K, IK, JK = ctx.matrix_to_upper_csr()
    assert np.array_equal(IK, golden[f"{s}.fd.IK"])
    assert np.array_equal(JK, golden[f"{s}.fd.JK"])
    assert_same(K, golden[f"{s}.fd.K_global"], "K_global")
    for n in ("dL", "csx", "csy", "csz", "dL_total", "TdL_total", "F"):
        assert_same(ctx.get_field(n), golden[f"{s}.fd.{n}"], f"side effect {n}")
    assert_same(ctx.get_field("Pin"), golden[f"{s}.fd.Pin"], "side effect Pin")


def test_predictor_bit_exact(ctx, golden):
    """computeBondForceGeneral(4, .): incremental elastic predictor with the stale cs* (constitutive.c:167-225)"""
    put_state(ctx, golden, "s1.fd")
    ctx.set_field("xyz", golden["s1.bc.xyz"])
    ctx.set_field("xyz_temp", golden["s1.pre.xyz"])
    ctx.set_field("F_temp", golden["s1.pre.F"])
    ctx.bond_force(4)
    for n in ("ddL", "F", "bond_stress"):
        assert_same(ctx.get_field(n), golden[f"s1.pred.{n}"], n)
    for n in ("ddL_total", "TddL_total", "stress_tensor", "J2_stresseq", "J2_stressm", "J2_triaxiality"):
        assert_same(ctx.get_field(n), golden[f"s1.pred.{n}"], n)
    assert_same(ctx.get_field("Pin"), golden["s1.pred.Pin"], "Pin")


def test_update_rr(ctx, golden):
    """updateRR(), stiffness.c:519-534 + the two dnrm2 of lpmc_project.c:412-413"""
    ctx.set_field("Pin", golden["s1.pred.Pin"])
    ctx.set_field("Pex", golden["s1.bc.Pex"])
    ctx.set_field("dispBC_index", golden["s1.bc.dispBC_index"])
    nr, nf = ctx.update_rr()
    assert_same(ctx.get_field("residual"), golden["s1.rr.residual"], "residual")
    assert nr == pytest.approx(golden["s1.rr.norms"][0], rel=1e-14)
    assert nf == pytest.approx(golden["s1.rr.norms"][1], rel=1e-14, abs=1e-300)


@pytest.mark.parametrize("tag,prev", [("s1.n0", "s1.pred"), ("s1.n1", "s1.n0.bf"), ("s1.n2", "s1.n1.bf")])
def test_j2_bond_force_bit_exact(ctx, golden, tag, prev):
    """computeBondForceGeneral(0, .) = computeBondForceJ2mixedLinear3D for every particle + computeStress +
    switchStateV(2) (constitutive.c:466-686): each particle evaluated once instead of 19x."""
    put_state(ctx, golden, prev)          # whatever the previous call left (stale triaxiality, pl_flag, ...)
    ctx.switch_state(0)                   # lpmc_project.c:428 (slot [1] is still the zero state in step 1)
    ctx.set_field("xyz", golden[f"{tag}.xyz"])
    ctx.bond_force(0)
    t = f"{tag}.bf"
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "dL_total", "TdL_total", "stress_tensor",
              "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "pl_flag"):
        assert_same(ctx.get_field(n), golden[f"{t}.{n}"], n)
    assert_same(ctx.get_field("Pin"), golden[f"{t}.Pin"], "Pin")
    assert_same(get_slots(ctx, "dLp", 3), golden[f"{t}.dLp"], "dLp")
    assert_same(get_slots(ctx, "J2_beta", 3), golden[f"{t}.J2_beta"], "J2_beta")
    assert_same(get_slots(ctx, "J2_alpha", 3), golden[f"{t}.J2_alpha"], "J2_alpha")
    assert int(golden[f"{t}.pl_flag"].sum()) > 0     # the case really is plastic


def test_j2_second_step_with_history(ctx, golden):
    """step 2, iteration 0: starts from the committed plastic state of step 1 (slot [1]) and damage_w < 1"""
    put_state(ctx, golden, "s2.pred")
    put_slots(ctx, "dLp", golden["s1.commit.dLp"])
    put_slots(ctx, "J2_alpha", golden["s1.commit.J2_alpha"])
    ctx.switch_state(0)
    ctx.set_field("xyz", golden["s2.n0.xyz"])
    ctx.bond_force(0)
    for n in ("F", "dL", "ddLp", "stress_tensor", "J2_dlambda"):
        assert_same(ctx.get_field(n), golden[f"s2.n0.bf.{n}"], n)
    assert_same(ctx.get_field("Pin"), golden["s2.n0.bf.Pin"], "Pin")
    assert_same(get_slots(ctx, "dLp", 3), golden["s2.n0.bf.dLp"], "dLp")
    assert golden["s2.pred.damage_w"].min() < 1.0


def test_elastic_law_matches_fd_base(ctx, golden):
    """plmode 6 on the unloaded lattice gives zero force; on a strained one Pin sums to ~0 (Newton's third law)"""
    ctx.bond_force(6)
    assert np.abs(ctx.get_field("F")).max() == 0.0
    xyz = golden["setup.xyz"].copy()
    xyz[:, 2] *= 1.001
    ctx.set_field("xyz", xyz)
    ctx.bond_force(6)
    Pin = ctx.get_field("Pin").reshape(-1, 3)
    assert np.abs(Pin.sum(axis=0)).max() < 1e-9 * np.abs(Pin).max()
    assert np.abs(ctx.get_field("F")).max() > 0


def test_update_crack_bit_exact(ctx, golden):
    """updateCrack(), constitutive.c:1399-1434 (scales F by damage_w a second time, rebuilds Pin)"""
    put_state(ctx, golden, "s1.dam")
    ctx.update_crack()
    assert_same(ctx.get_field("F"), golden["s1.crack.F"], "F")
    assert_same(ctx.get_field("Pin"), golden["s1.crack.Pin"], "Pin")
    assert_same(ctx.get_field("nb"), golden["s1.crack.nb"], "nb")
    assert_same(ctx.get_field("damage_visual"), golden["s1.crack.damage_visual"], "damage_visual")


def test_switch_state_semantics(ctx):
    """switchStateV: 0: [0]:=[1]; 1: [1]:=[0]; 2: [0]:=[2] without damage_* (constitutive.c:10-85)"""
    N, nn = ctx.N, ctx.nn
    rng = np.random.default_rng(20240607)
    a, b, d = rng.standard_normal((3, N, nn))
    ctx.set_field("dLp0", a); ctx.set_field("dLp1", b); ctx.set_field("dLp2", d)
    ctx.set_field("damage_D0", a); ctx.set_field("damage_D1", b)
    ctx.switch_state(2)
    assert_same(ctx.get_field("dLp0"), d); assert_same(ctx.get_field("damage_D0"), a)
    ctx.switch_state(1)
    assert_same(ctx.get_field("dLp1"), d); assert_same(ctx.get_field("damage_D1"), a)
    ctx.set_field("dLp1", b)
    ctx.switch_state(0)
    assert_same(ctx.get_field("dLp0"), b)


@pytest.mark.parametrize("tag,t", [("s1.n0", 1), ("s1.n1", 1), ("s1.n2", 1), ("s2.n0", -1), ("s2.n1", -1), ("s2.n2", -1)])
def test_j2_energy_law_bit_exact(lpm, tag, t):
    """computeBondForceGeneral(3, t) = computeBondForceJ2energyReturnMap (constitutive.c:286-463):
    bisection return map on the distortional energy; every recorded call of tests/golden/sc6_j2energy.npz is
    replayed from its complete input state.  Step 2 has broken bonds (nb < nb_initial) and t = -1."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2energy.npz")
    c = make_ctx(lpm, g)
    pre = f"{tag}.pre"
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w", "dL_total", "TdL_total",
              "stress_tensor", "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "xyz", "Pin", "pl_flag", "nb"):
        c.set_field(n, g[f"{pre}.{n}"])
    put_slots(c, "dLp", g[f"{pre}.dLp"])
    put_slots(c, "damage_D", g[f"{pre}.damage_D"])
    put_slots(c, "J2_alpha", g[f"{pre}.J2_alpha"])
    put_slots(c, "J2_beta_eq", g[f"{pre}.J2_beta_eq"])
    c.bond_force(3, t)
    bf = f"{tag}.bf"
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "dL_total", "TdL_total", "stress_tensor", "J2_dlambda",
              "J2_stresseq", "J2_stressm", "J2_triaxiality", "pl_flag", "Pin"):
        assert_same(c.get_field(n), g[f"{bf}.{n}"], n)
    assert_same(get_slots(c, "dLp", 3), g[f"{bf}.dLp"], "dLp")
    assert_same(get_slots(c, "J2_alpha", 3), g[f"{bf}.J2_alpha"], "J2_alpha")
    assert_same(get_slots(c, "J2_beta_eq", 3), g[f"{bf}.J2_beta_eq"], "J2_beta_eq")
    if tag != "s2.n0":
        assert (g[f"{bf}.J2_dlambda"] > 0).sum() > 50      # the bisection really ran
    if tag.startswith("s2"):
        assert (g[f"{pre}.nb"] < g["setup.nb_initial"]).any()
    c.close()


@pytest.mark.parametrize("step", ["s1", "s2"])
def test_compute_strain_bit_exact(ctx, golden, step):
    """computeStrain(), lpm_basic.c:127-249: weighted least squares + 6x6 LU per particle"""
    ctx.set_field("dL", golden[f"{step}.strain.dL"])
    ctx.set_field("strain_tensor", np.zeros((216, 6)) if step == "s1" else golden["s1.strain.strain_tensor"])
    ctx.compute_strain()
    assert_same(ctx.get_field("strain_tensor"), golden[f"{step}.strain.strain_tensor"], "strain_tensor")
    assert np.abs(golden[f"{step}.strain.strain_tensor"]).max() > 1e-4


def _j2iso_ctx(lpm, g, pre):
    c = make_ctx(lpm, g)
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w", "dL_total", "TdL_total",
              "stress_tensor", "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "xyz", "Pin", "pl_flag", "nb"):
        c.set_field(n, g[f"{pre}.{n}"])
    put_slots(c, "dLp", g[f"{pre}.dLp"])
    put_slots(c, "damage_D", g[f"{pre}.damage_D"])
    put_slots(c, "J2_alpha", g[f"{pre}.J2_alpha"])
    put_slots(c, "J2_beta", g[f"{pre}.J2_beta"])
    put_slots(c, "J2_beta_eq", g[f"{pre}.J2_beta_eq"])
    put_slots(c, "damage_local", g[f"{pre}.damage_local"])
    return c


@pytest.mark.parametrize("tag", ["s1.n0", "s1.n1", "s2.n0", "s3.n0", "s3.n1"])
def test_j2_nonlinear_iso_law_serial_semantics(lpm, tag):
    """computeBondForceGeneral(5, .) = computeBondForceJ2nonlinearIso (constitutive.c:689-863): the law
    updates slot [0] in place 1 + nb times per particle in the order of the serial loop; the per-particle trajectory
    kernel reproduces that exactly (tests/golden/sc6_j2iso.npz, single-threaded reference).  Step 3 has broken bonds.
    exp() of SY(x) is the only non-IEEE operation: it only steers bisection decisions, so results stay bit-exact
    unless a decision sits within an ulp of a tie."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2iso.npz")
    c = _j2iso_ctx(lpm, g, f"{tag}.pre")
    c.bond_force(5, 1)
    bf = f"{tag}.bf"
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "dL_total", "TdL_total", "stress_tensor", "J2_dlambda",
              "J2_stresseq", "J2_stressm", "J2_triaxiality", "Pin"):
        assert_same(c.get_field(n), g[f"{bf}.{n}"], n)
    assert_same(get_slots(c, "dLp", 3), g[f"{bf}.dLp"], "dLp")
    assert_same(get_slots(c, "J2_alpha", 3), g[f"{bf}.J2_alpha"], "J2_alpha")
    assert_same(get_slots(c, "J2_beta", 3), g[f"{bf}.J2_beta"], "J2_beta")
    if tag.startswith("s3"):
        assert (g[f"{tag}.pre.nb"] < g["setup.nb_initial"]).any()
    c.close()


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_local_bondwise_damage_bit_exact(lpm, step):
    """updateDamageGeneral(., ., 5) = updateDuctileDamageBwiseLocal (constitutive.c:1607-1695); step 2 breaks 18 bonds"""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2iso.npz")
    c = _j2iso_ctx(lpm, g, f"{step}.dam.pre")
    broken, pairs = c.update_damage(5)
    assert broken == int(g[f"{step}.dam.broken"][0])
    for n in ("damage_broken", "damage_w", "nb"):
        assert_same(c.get_field(n), g[f"{step}.dam.{n}"], n)
    assert_same(get_slots(c, "damage_D", 2), g[f"{step}.dam.damage_D"], "damage_D")
    assert_same(get_slots(c, "damage_local", 2), g[f"{step}.dam.damage_local"], "damage_local")
    if step == "s2":
        assert broken == 18 and len(pairs) == 18 and all(i < j for i, j in pairs)
    c.close()
