"""GPU parity: crystal plasticity (plmode 1) -- computeCab and computeBondForceCPMiehe (src/constitutive.c:866-1396,
1864-1917) on the FCC / Al case of examples/FCC_Al_R0.3_001_tension.c (BASELINE config 4; the shipped example does not
compile against the reference's current sources, its library functions do -- tests/golden/make_golden_cp.py drives
them).  Cab is + - * / only: bit-exact.  The return map uses pow / cosh / tanh (<= 1-2 ulp from glibc) and a 24x24 LU:
compared at 1e-9 relative (north_star's tolerance), discrete outputs (active sets) exactly."""
import numpy as np
import pytest

from helpers import assert_same, put_slots, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def gcp():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "fcc_cp.npz")


def make_cp_ctx(lpm, g):
    par = {str(k): float(v) for k, v in zip(g["param_names"], g["params"])}
    N, nn = g["setup.neighbors"].shape
    c = lpm.Context(N, 3, 3, nn, g["setup.conn"].shape[1])
    c.set_params(**{k: v for k, v in par.items() if k != "nslipSys"})
    c.set_field("xyz", g["setup.xyz"])
    c.set_field("xyz_initial", g["setup.xyz"])
    c.build_topology(par["neighbor1_cutoff"], par["neighbor2_cutoff"])
    c.set_field("type", g["setup.type"])
    c.calc_kntv(g["setup.Ce"])
    c.compute_dl()
    c.set_schmid_tensor(g["setup.schmid_tensor"])
    put_slots(c, "cp_gy", g["setup.cp_gy"])
    return c, par


def test_fcc_topology_and_kntv(lpm, gcp):
    """FCC lattice: 12 + 6 neighbours, stiffness.c:207-235"""
    c, _ = make_cp_ctx(lpm, gcp)
    assert_same(c.get_field("neighbors"), gcp["setup.neighbors"], "neighbors")
    assert_same(c.get_field("nsign"), gcp["setup.nsign"], "nsign")
    assert np.array_equal(c.k_pointer(), gcp["setup.K_pointer"])
    assert_same(c.get_field("Kn"), gcp["setup.Kn"], "Kn")
    assert_same(c.get_field("Tv"), gcp["setup.Tv"], "Tv")
    assert_same(c.get_field("distance"), gcp["setup.distance"], "distance")
    c.close()


def test_compute_cab_bit_exact(lpm, gcp):
    c, _ = make_cp_ctx(lpm, gcp)
    c.compute_cab()
    assert_same(c.get_field("cp_Cab"), gcp["setup.cp_Cab"], "cp_Cab")
    c.close()


@pytest.mark.parametrize("tag,prev", [("s1.n0", "s1.pred"), ("s1.n1", "s1.n0.bf"), ("s1.n2", "s1.n1.bf")])
def test_cp_bond_force(lpm, gcp, tag, prev):
    g = gcp
    c, _ = make_cp_ctx(lpm, g)
    c.set_field("cp_Cab", g["setup.cp_Cab"])
    for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "damage_broken", "damage_w", "J2_triaxiality", "pl_flag"):
        c.set_field(n, g[f"{prev}.{n}"])
    put_slots(c, "dLp", g[f"{prev}.dLp"])
    put_slots(c, "cp_gy", g[f"{prev}.cp_gy"])
    put_slots(c, "cp_A_single", g[f"{prev}.cp_A_single"])
    put_slots(c, "cp_A", g[f"{prev}.cp_A"])
    c.switch_state(0)
    c.set_field("xyz", g[f"{tag}.xyz"])
    c.bond_force(1)
    t = f"{tag}.bf"
    assert np.array_equal(c.get_field("cp_Jact"), g[f"{t}.cp_Jact"]), "active slip systems differ"
    assert int(g[f"{t}.cp_Jact"].sum()) > 0
    for n in ("F", "dL", "ddLp", "stress_tensor", "cp_RSS", "cp_dgy", "cp_dA", "cp_dA_single", "bond_stress"):
        assert rel_err(c.get_field(n), g[f"{t}.{n}"]) <= TOL, n
    assert rel_err(c.get_field("Pin"), g[f"{t}.Pin"]) <= 100 * TOL   # sums of cancelling bond forces
    for n, k in (("dLp", 3), ("cp_gy", 3), ("cp_A_single", 3), ("cp_A", 3)):
        got = np.stack([c.get_field(f"{n}{s}") for s in range(k)], axis=-1)
        assert rel_err(got, g[f"{t}.{n}"]) <= TOL, n
    assert np.array_equal(c.get_field("pl_flag"), g[f"{t}.pl_flag"])
    c.close()


@pytest.mark.parametrize("tag,prev", [("s1.n0", "s1.pred"), ("s1.n2", "s1.n1.bf")])
def test_warp_kernel_equals_thread_kernel(lpm, gcp, tag, prev):
    """cp_miehe_warp_kernel (one warp per particle, lane = slip system, LU across the lanes; the default for <= 24 slip
    systems) against cp_miehe_kernel (one thread per particle; param cp_warp = 0): every sum is accumulated in the same
    order and the LU applies the same operations to every element, so all outputs must agree BIT FOR BIT"""
    from helpers import assert_same
    g = gcp
    out = {}
    for warp in (0, 1):
        c, _ = make_cp_ctx(lpm, g)
        c.set_params(cp_warp=float(warp))
        c.set_field("cp_Cab", g["setup.cp_Cab"])
        for n in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "damage_broken", "damage_w", "J2_triaxiality", "pl_flag"):
            c.set_field(n, g[f"{prev}.{n}"])
        put_slots(c, "dLp", g[f"{prev}.dLp"])
        put_slots(c, "cp_gy", g[f"{prev}.cp_gy"])
        put_slots(c, "cp_A_single", g[f"{prev}.cp_A_single"])
        put_slots(c, "cp_A", g[f"{prev}.cp_A"])
        c.switch_state(0)
        c.set_field("xyz", g[f"{tag}.xyz"])
        c.bond_force(1)
        out[warp] = {n: c.get_field(n) for n in ("F", "Pin", "dL", "ddLp", "stress_tensor", "cp_RSS", "cp_dgy", "cp_dA", "cp_dA_single",
                                                  "cp_Jact", "pl_flag", "dLp2", "cp_gy2", "cp_A2", "cp_A_single2")}
        c.close()
    assert int(out[0]["cp_Jact"].sum()) > 0
    for n in out[0]:
        assert_same(out[1][n], out[0][n], f"{n} (warp kernel vs thread kernel)")


def test_cp_two_load_steps(lpm, gcp):
    """whole load steps, device resident (driver.load_step with plmode 1) vs the reference's run"""
    g = gcp
    c, _ = make_cp_ctx(lpm, g)
    c.compute_cab()
    dbp = [(1, "z", 0.0), (2, "x", 0.0), (2, "z", 0.0), (3, "y", 0.0), (3, "z", 0.0), (4, "x", 0.0), (4, "y", 0.0), (4, "z", 0.0),
           (5, "z", -2.0e-3)]
    for step in (1, 2):
        log = lpm.driver.load_step(c, 1, dbp, [])
        assert log.newton_iterations == int(g["newton_counts"][step - 1])
        s = f"s{step}.end"
        u, u_ref = c.get_field("xyz") - g["setup.xyz"], g[f"{s}.xyz"] - g["setup.xyz"]
        assert rel_err(u, u_ref) <= TOL
        assert rel_err(c.get_field("F"), g[f"{s}.F"]) <= 10 * TOL
        assert rel_err(c.get_field("stress_tensor"), g[f"{s}.stress_tensor"]) <= 10 * TOL
        assert rel_err(c.get_field("cp_A0"), g[f"{s}.cp_A"][:, 0]) <= 10 * TOL
        assert np.array_equal(c.get_field("cp_Jact"), g[f"{s}.cp_Jact"])
    c.close()
