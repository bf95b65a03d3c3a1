"""CPU: pin the C restatement (oracle/lpm_oracle.c) against the reference's own functions.

The golden vectors were produced by the unmodified reference sources (oracle/_ref); the restatement must
reproduce them BIT FOR BIT through a whole plastic load step (set-up -> FD tangent -> predictor -> residual ->
BC-modified tangent -> CG -> J2 bond force -> ... -> damage -> crack update)."""
import numpy as np
import pytest

from helpers import assert_same, params_from_golden


@pytest.fixture(scope="module")
def port(golden):
    from oracle import port as P
    if not P.available():
        pytest.skip("oracle/liblpm_oracle.so not built")
    g = golden
    par = params_from_golden(g)
    p = P.Port(g["setup.xyz"], radius=par["radius"], particle_volume=par["particle_volume"])
    p.search_neighbors(par["neighbor1_cutoff"], par["neighbor2_cutoff"])
    p.type[:] = g["setup.type"]
    p.sigmay[:] = g["setup.sigmay"]
    p.Kn[:], p.Tv[:] = g["setup.Kn"], g["setup.Tv"]
    p.Ce = np.ascontiguousarray(g["setup.Ce"])
    p.J2_H, p.J2_xi = par["J2_H"], par["J2_xi"]
    p.par = par
    return p


def test_port_whole_first_load_step_bit_exact(port, golden):
    g, p = golden, port
    # neighbor.c
    assert_same(p.neighbors, g["setup.neighbors"], "neighbors"); assert_same(p.nsign, g["setup.nsign"], "nsign")
    assert_same(p.conn, g["setup.conn"], "conn"); assert_same(p.nb_conn, g["setup.nb_conn"], "nb_conn")
    assert_same(np.stack([p.kp0, p.kp1], 1), g["setup.K_pointer"].astype(np.int64), "K_pointer")
    for n in ("distance_initial", "csx_initial", "csy_initial", "csz_initial"):
        assert_same(getattr(p, n), g[f"setup.{n}"], n)
    # computedL
    p.computedL()
    for n in ("distance", "dL", "dL_total", "TdL_total", "csx"):
        assert_same(getattr(p, n), g[f"setup.{n}"], n)
    # FD tangent + side effects
    p.xyz_temp[:] = p.xyz; p.F_temp[:] = p.F
    p.calcStiffnessFiniteDifference()
    assert_same(p.K_global, g["s1.fd.K_global"], "K_global"); assert_same(p.IK, g["s1.fd.IK"], "IK"); assert_same(p.JK, g["s1.fd.JK"], "JK")
    for n in ("dL", "csx", "csz", "dL_total", "TdL_total", "F", "Pin"):
        assert_same(getattr(p, n), g[f"s1.fd.{n}"], f"fd side effect {n}")
    # BCs (host, boundary.c:12-70) taken from the fixture; predictor; residual
    p.xyz[:] = g["s1.bc.xyz"]; p.Pex[:] = g["s1.bc.Pex"]; p.dispBC_index[:] = g["s1.bc.dispBC_index"]
    p.computeBondForceGeneral(4)
    for n in ("ddL", "F", "Pin", "stress_tensor", "bond_stress"):
        assert_same(getattr(p, n), g[f"s1.pred.{n}"], f"predictor {n}")
    p.updateRR()
    assert_same(p.residual, g["s1.rr.residual"], "residual")
    # three Newton iterations
    for it in range(3):
        t = f"s1.n{it}"
        p.switchStateV(0)
        p.setDispBC_stiffnessUpdate()
        assert_same(p.K_global, g[f"{t}.K_bc"], "K_bc"); assert_same(p.residual, g[f"{t}.rhs"], "rhs")
        iters = p.solverCG()
        assert iters == int(g[f"{t}.cg_iters"][0])
        assert_same(p.disp, g[f"{t}.disp"], "disp"); assert_same(p.xyz, g[f"{t}.xyz"], "xyz")
        p.computeBondForceGeneral(0)
        for n in ("dL", "dL_ave", "ddLp", "F", "Pin", "stress_tensor", "J2_dlambda", "J2_triaxiality", "pl_flag", "bond_stress"):
            assert_same(getattr(p, n), g[f"{t}.bf.{n}"], f"{t} {n}")
        assert_same(np.stack([p.dLp0, p.dLp1, p.dLp2], -1), g[f"{t}.bf.dLp"], "dLp")
        assert_same(np.stack(p.J2_alpha, -1), g[f"{t}.bf.J2_alpha"], "J2_alpha")
        p.updateRR()
        assert_same(p.residual, g[f"{t}.residual"], "residual")


def test_port_damage_and_crack(port, golden):
    from helpers import BOND
    g, p = golden, port
    for n in ("F", "csx", "csy", "csz", "damage_broken", "damage_w"):
        getattr(p, n)[:] = g[f"s1.n2.bf.{n}"]
    p.J2_dlambda[:], p.J2_triaxiality[:] = g["s1.dam.J2_dlambda"], g["s1.dam.J2_triaxiality"]
    p.damage_nonlocal0[:] = 0.0
    p.damage_D0[:] = 0.0
    k = p.updateDamageNonlocal(p.par["damage_L"], p.par["damage_threshold"], p.par["damagec_A"])
    assert k == int(g["s1.dam.broken"][0])
    assert_same(p.damage_nonlocal0, g["s1.dam.damage_nonlocal"][:, 0], "damage_nonlocal")
    assert_same(p.damage_w, g["s1.dam.damage_w"], "damage_w")
    for n in ("F", "csx", "csy", "csz"):
        getattr(p, n)[:] = g[f"s1.dam.{n}"]
    p.updateCrack()
    assert_same(p.F, g["s1.crack.F"], "F"); assert_same(p.Pin, g["s1.crack.Pin"], "Pin")
    assert_same(p.nb, g["s1.crack.nb"], "nb")
