"""CPU: pin the C restatement (oracle/lpm_oracle.c) against the reference's own functions.

The golden vectors were produced by the unmodified reference sources (oracle/_ref); the restatement must
reproduce them BIT FOR BIT through a whole plastic load step (set-up -> FD tangent -> predictor -> residual ->
BC-modified tangent -> CG -> J2 bond force -> ... -> damage -> crack update)."""
import os

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


# ----------------------------------------------------------------------------------------------------------------
# rows a8 / f4: the alternative J2 laws, the local bond-wise damage and computeStrain, restated in lpm_oracle.c as
# the reference's literal serial loops, against the fixtures produced by the reference itself
def _lib():
    import ctypes as C
    from oracle import port as P
    if not P.available():
        pytest.skip("oracle/liblpm_oracle.so not built")
    lib = C.CDLL(str(P.SO))
    lib.oracle_damage_local_bondwise.restype = C.c_int
    return lib, C


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a):
    import ctypes as C
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("tag,t", [("s1.n0", 1), ("s1.n2", 1), ("s2.n0", -1), ("s2.n2", -1)])
def test_port_j2_energy_law_bit_exact(tag, t):
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2energy.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    pre, bf = f"{tag}.pre", f"{tag}.bf"
    f8, i4 = np.float64, np.int32
    a = {k: _c(g[f"{pre}.{k}"], f8) for k in ("xyz", "damage_broken", "ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin",
                                              "J2_dlambda")}
    dLp, alpha, beq, dD = g[f"{pre}.dLp"], g[f"{pre}.J2_alpha"], g[f"{pre}.J2_beta_eq"], g[f"{pre}.damage_D"]
    dLp0, dLp2 = _c(dLp[..., 0], f8), _c(dLp[..., 2], f8)
    a0, a2, b0, b2 = _c(alpha[:, 0], f8), _c(alpha[:, 2], f8), _c(beq[:, 0], f8), _c(beq[:, 2], f8)
    plf, nb = _c(g[f"{pre}.pl_flag"], i4), _c(g[f"{pre}.nb"], i4)
    cst = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial", "type")}
    cf = {k: _c(g[f"setup.{k}"], f8) for k in ("Ce", "sigmay", "distance_initial", "Kn", "Tv")}
    lib.oracle_j2_energy_force(C.c_int(N), C.c_int(nn), C.c_double(par["particle_volume"]), C.c_double(par["radius"]), C.c_double(par["J2_H"]),
                               C.c_double(par["J2_xi"]), C.c_int(t), _ptr(cf["Ce"]), _ptr(cst["type"]), _ptr(cf["sigmay"]), _ptr(a["xyz"]),
                               _ptr(cst["neighbors"]), _ptr(cst["nsign"]), _ptr(cst["nb_initial"]), _ptr(nb), _ptr(cf["distance_initial"]),
                               _ptr(cf["Kn"]), _ptr(cf["Tv"]), _ptr(a["damage_broken"]), _ptr(_c(dD[..., 0], f8)), _ptr(dLp0), _ptr(b0), _ptr(a0),
                               _ptr(dLp2), _ptr(b2), _ptr(a2), _ptr(a["J2_dlambda"]), _ptr(plf), _ptr(a["ddLp"]), _ptr(a["dL"]), _ptr(a["dL_ave"]),
                               _ptr(a["dL_total"]), _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["F"]), _ptr(a["Pin"]))
    for k in ("ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin", "J2_dlambda"):
        assert_same(a[k], g[f"{bf}.{k}"], k)
    assert_same(plf, g[f"{bf}.pl_flag"], "pl_flag")
    assert_same(dLp2, g[f"{bf}.dLp"][..., 2], "dLp[2]")
    assert_same(a2, g[f"{bf}.J2_alpha"][:, 2], "J2_alpha[2]")
    assert_same(b2, g[f"{bf}.J2_beta_eq"][:, 2], "J2_beta_eq[2]")


@pytest.mark.parametrize("tag", ["s1.n0", "s2.n0", "s3.n0", "s3.n1"])
def test_port_j2_iso_law_bit_exact(tag):
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2iso.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    pre, bf = f"{tag}.pre", f"{tag}.bf"
    f8, i4 = np.float64, np.int32
    a = {k: _c(g[f"{pre}.{k}"], f8) for k in ("xyz", "damage_broken", "damage_w", "ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz",
                                              "F", "Pin", "J2_dlambda")}
    dLp0 = _c(g[f"{pre}.dLp"][..., 0], f8)
    beta0 = _c(g[f"{pre}.J2_beta"][..., 0], f8)
    alpha0 = _c(g[f"{pre}.J2_alpha"][:, 0], f8)
    nb = _c(g[f"{pre}.nb"], i4)
    cst = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial", "type")}
    cf = {k: _c(g[f"setup.{k}"], f8) for k in ("Ce", "distance_initial", "Kn", "Tv")}
    lib.oracle_j2_iso_force(C.c_int(N), C.c_int(nn), C.c_double(par["particle_volume"]), C.c_double(par["J2_C"]), _ptr(cf["Ce"]), _ptr(cst["type"]),
                            _ptr(a["xyz"]), _ptr(cst["neighbors"]), _ptr(cst["nsign"]), _ptr(cst["nb_initial"]), _ptr(nb),
                            _ptr(cf["distance_initial"]), _ptr(cf["Kn"]), _ptr(cf["Tv"]), _ptr(a["damage_broken"]), _ptr(a["damage_w"]), _ptr(dLp0),
                            _ptr(beta0), _ptr(alpha0), _ptr(a["J2_dlambda"]), _ptr(a["ddLp"]), _ptr(a["dL"]), _ptr(a["dL_ave"]), _ptr(a["dL_total"]),
                            _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["F"]), _ptr(a["Pin"]))
    for k in ("ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin", "J2_dlambda"):
        assert_same(a[k], g[f"{bf}.{k}"], k)
    assert np.abs(dLp0).max() > 1e-4      # plastic stretch accumulated in place (the reference wipes it afterwards)


def _params2(g, pre):
    return {str(k): float(v) for k, v in zip(g[f"{pre}.param_names"], g[f"{pre}.params"])}


@pytest.mark.parametrize("tag", ["e.s1", "e.s2"])
def test_port_j2_energy_law_called_per_particle_bit_exact(tag):
    """computeBondForceJ2energyReturnMap(ii, t) called on its own (constitutive.h:21), five particles in sequence, against
    tests/golden/sc6_particle2.npz: every array the call may write, after every call (the star's rows change, the others
    do not; across a broken bond the force pass reads the partner's stale rows)"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_particle2.npz")
    par = _params2(g, "e")
    N, nn = g["e.setup.neighbors"].shape
    pre = f"{tag}.pre"
    t = int(g[f"{tag}.t"][0])
    f8, i4 = np.float64, np.int32
    a = {k: _c(g[f"{pre}.{k}"], f8) for k in ("xyz", "damage_broken", "ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin",
                                              "J2_dlambda")}
    dLp, alpha, beq, dD = g[f"{pre}.dLp"], g[f"{pre}.J2_alpha"], g[f"{pre}.J2_beta_eq"], g[f"{pre}.damage_D"]
    dLp0, dLp2 = _c(dLp[..., 0], f8), _c(dLp[..., 2], f8)
    a0, a2, b0, b2 = _c(alpha[:, 0], f8), _c(alpha[:, 2], f8), _c(beq[:, 0], f8), _c(beq[:, 2], f8)
    plf, nb = _c(g[f"{pre}.pl_flag"], i4), _c(g[f"{pre}.nb"], i4)
    cst = {k: _c(g[f"e.setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial", "type")}
    cf = {k: _c(g[f"e.setup.{k}"], f8) for k in ("Ce", "sigmay", "distance_initial", "Kn", "Tv")}
    dD0 = _c(dD[..., 0], f8)
    for k, ii in enumerate(g[f"{tag}.particles"]):
        ii = int(ii)
        lib.oracle_j2_energy_force_range(C.c_int(ii), C.c_int(ii + 1), C.c_int(N), C.c_int(nn), C.c_double(par["particle_volume"]),
                                         C.c_double(par["radius"]), C.c_double(par["J2_H"]), C.c_double(par["J2_xi"]), C.c_int(t), _ptr(cf["Ce"]),
                                         _ptr(cst["type"]), _ptr(cf["sigmay"]), _ptr(a["xyz"]), _ptr(cst["neighbors"]), _ptr(cst["nsign"]),
                                         _ptr(cst["nb_initial"]), _ptr(nb), _ptr(cf["distance_initial"]), _ptr(cf["Kn"]), _ptr(cf["Tv"]),
                                         _ptr(a["damage_broken"]), _ptr(dD0), _ptr(dLp0), _ptr(b0), _ptr(a0), _ptr(dLp2), _ptr(b2), _ptr(a2),
                                         _ptr(a["J2_dlambda"]), _ptr(plf), _ptr(a["ddLp"]), _ptr(a["dL"]), _ptr(a["dL_ave"]), _ptr(a["dL_total"]),
                                         _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["F"]), _ptr(a["Pin"]))
        for n in ("ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin", "J2_dlambda"):
            assert_same(a[n], g[f"{tag}.c{k}.{n}"], f"call {k} (particle {ii}): {n}")
        assert_same(plf, g[f"{tag}.c{k}.pl_flag"], "pl_flag")
        assert_same(dLp2, g[f"{tag}.c{k}.dLp2"], "dLp[2]")
        assert_same(a2, g[f"{tag}.c{k}.J2_alpha2"], "J2_alpha[2]")
        assert_same(b2, g[f"{tag}.c{k}.J2_beta_eq2"], "J2_beta_eq[2]")
    assert (a["dL"] != g[f"{pre}.dL"]).any()


@pytest.mark.parametrize("tag", ["i.s1", "i.s2"])
def test_port_j2_iso_law_called_per_particle_bit_exact(tag):
    """computeBondForceJ2nonlinearIso(ii) called on its own (constitutive.h:18): the star's plastic state advances in
    place (slot [0]), star members keep their trial forces in F, ii gets its final F / Pin and a zeroed stress row"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_particle2.npz")
    par = _params2(g, "i")
    N, nn = g["i.setup.neighbors"].shape
    pre = f"{tag}.pre"
    f8, i4 = np.float64, np.int32
    a = {k: _c(g[f"{pre}.{k}"], f8) for k in ("xyz", "damage_broken", "damage_w", "ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz",
                                              "F", "Pin", "J2_dlambda", "stress_tensor")}
    dLp0 = _c(g[f"{pre}.dLp"][..., 0], f8)
    beta0 = _c(g[f"{pre}.J2_beta"][..., 0], f8)
    alpha0 = _c(g[f"{pre}.J2_alpha"][:, 0], f8)
    nb = _c(g[f"{pre}.nb"], i4)
    cst = {k: _c(g[f"i.setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial", "type")}
    cf = {k: _c(g[f"i.setup.{k}"], f8) for k in ("Ce", "distance_initial", "Kn", "Tv")}
    for k, ii in enumerate(g[f"{tag}.particles"]):
        ii = int(ii)
        lib.oracle_j2_iso_force_range(C.c_int(ii), C.c_int(ii + 1), C.c_int(N), C.c_int(nn), C.c_double(par["particle_volume"]),
                                      C.c_double(par["J2_C"]), _ptr(cf["Ce"]), _ptr(cst["type"]), _ptr(a["xyz"]), _ptr(cst["neighbors"]),
                                      _ptr(cst["nsign"]), _ptr(cst["nb_initial"]), _ptr(nb), _ptr(cf["distance_initial"]), _ptr(cf["Kn"]),
                                      _ptr(cf["Tv"]), _ptr(a["damage_broken"]), _ptr(a["damage_w"]), _ptr(dLp0), _ptr(beta0), _ptr(alpha0),
                                      _ptr(a["J2_dlambda"]), _ptr(a["ddLp"]), _ptr(a["dL"]), _ptr(a["dL_ave"]), _ptr(a["dL_total"]),
                                      _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["F"]), _ptr(a["Pin"]))
        a["stress_tensor"][ii] = 0.0      # memset(stress_tensor[ii], 0, ...), constitutive.c:834 (the dispatcher recomputes the stress)
        for n in ("ddLp", "dL", "dL_ave", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin", "J2_dlambda", "stress_tensor"):
            assert_same(a[n], g[f"{tag}.c{k}.{n}"], f"call {k} (particle {ii}): {n}")
        assert_same(dLp0, g[f"{tag}.c{k}.dLp0"], "dLp[0]")
        assert_same(alpha0, g[f"{tag}.c{k}.J2_alpha0"], "J2_alpha[0]")
        assert_same(beta0, g[f"{tag}.c{k}.J2_beta0"], "J2_beta[0]")
    assert (dLp0 != g[f"{pre}.dLp"][..., 0]).any()


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_port_local_bondwise_damage_bit_exact(step):
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_j2iso.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    pre, out = f"{step}.dam.pre", f"{step}.dam"
    f8, i4 = np.float64, np.int32
    dloc0 = _c(g[f"{pre}.damage_local"][:, 0], f8)
    broken, w, dD0 = _c(g[f"{pre}.damage_broken"], f8), _c(g[f"{pre}.damage_w"], f8), _c(g[f"{pre}.damage_D"][..., 0], f8)
    nb = _c(g[f"{pre}.nb"], i4)
    pairs = np.full((64, 2), -1, i4)
    k = lib.oracle_damage_local_bondwise(C.c_int(N), C.c_int(nn), C.c_double(par["damage_threshold"]), C.c_double(par["damagec_A"]),
                                         _ptr(_c(g["setup.neighbors"], i4)), _ptr(_c(g["setup.nb_initial"], i4)),
                                         _ptr(_c(g[f"{pre}.J2_triaxiality"], f8)), _ptr(_c(g[f"{pre}.J2_dlambda"], f8)), _ptr(dloc0), _ptr(broken),
                                         _ptr(dD0), _ptr(w), _ptr(nb), _ptr(pairs), C.c_int(64))
    assert k == int(g[f"{out}.broken"][0])
    assert_same(dloc0, g[f"{out}.damage_local"][:, 0], "damage_local")
    assert_same(broken, g[f"{out}.damage_broken"], "damage_broken")
    assert_same(dD0, g[f"{out}.damage_D"][..., 0], "damage_D")
    assert_same(w, g[f"{out}.damage_w"], "damage_w")
    assert_same(nb, g[f"{out}.nb"], "nb")


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_port_particlewise_local_damage_bit_exact(step):
    """updateDuctileDamagePwiseLocal (constitutive.c:1529-1579) restated, against tests/golden/sc6_damage_variants.npz"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_damage_variants.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    pre, out = f"pwl.{step}.pre", f"pwl.{step}.post"
    f8, i4 = np.float64, np.int32
    dloc0 = _c(g[f"{pre}.damage_local"][:, 0], f8)
    broken, w, dD0 = _c(g[f"{pre}.damage_broken"], f8), _c(g[f"{pre}.damage_w"], f8), _c(g[f"{pre}.damage_D"][..., 0], f8)
    lst = np.full(64, -1, i4)
    k = lib.oracle_damage_local_particlewise(C.c_int(N), C.c_int(nn), C.c_double(par["damage_threshold"]), C.c_double(par["damagec_A"]),
                                             _ptr(_c(g["setup.neighbors"], i4)), _ptr(_c(g["setup.nb_initial"], i4)),
                                             _ptr(_c(g[f"{pre}.J2_triaxiality"], f8)), _ptr(_c(g[f"{pre}.J2_dlambda"], f8)), _ptr(dloc0),
                                             _ptr(broken), _ptr(dD0), _ptr(w), _ptr(lst), C.c_int(64))
    assert k == int(g[f"pwl.{step}.broken"][0])
    assert_same(dloc0, g[f"{out}.damage_local"][:, 0], "damage_local")
    assert_same(broken, g[f"{out}.damage_broken"], "damage_broken")
    assert_same(dD0, g[f"{out}.damage_D"][..., 0], "damage_D")
    assert_same(w, g[f"{out}.damage_w"], "damage_w")
    assert np.array_equal(lst[:k], np.flatnonzero((g[f"{out}.damage_local"][:, 0] == 1.0) & (g[f"{pre}.damage_local"][:, 0] != 1.0)))


@pytest.mark.parametrize("step", ["s1", "s2", "s3"])
def test_port_bondwise_nonlocal_damage_bit_exact(step):
    """updateDuctileDamageBwiseNonlocal (constitutive.c:1698-1753) restated, against tests/golden/sc6_damage_variants.npz"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_damage_variants.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    pre, out = f"bwn.{step}.pre", f"bwn.{step}.post"
    f8, i4 = np.float64, np.int32
    Dn = _c(g[f"{pre}.damage_nonlocal"][:, 0], f8)
    broken, w, dD0 = _c(g[f"{pre}.damage_broken"], f8), _c(g[f"{pre}.damage_w"], f8), _c(g[f"{pre}.damage_D"][..., 0], f8)
    pairs = np.full((512, 2), -1, i4)
    k = lib.oracle_damage_nonlocal_bondwise(C.c_int(N), C.c_int(nn), C.c_double(par["damage_L"]), C.c_double(par["damage_threshold"]),
                                            C.c_double(par["damagec_A"]), C.c_double(par["particle_volume"]),
                                            _ptr(_c(g["setup.neighbors"], i4)), _ptr(_c(g["setup.nb_initial"], i4)),
                                            _ptr(_c(g["setup.distance_initial"], f8)), _ptr(_c(g[f"{pre}.J2_dlambda"], f8)),
                                            _ptr(_c(g[f"{pre}.J2_triaxiality"], f8)), _ptr(Dn), _ptr(broken), _ptr(dD0), _ptr(w), _ptr(pairs),
                                            C.c_int(512))
    assert k == int(g[f"bwn.{step}.broken"][0])
    assert_same(Dn, g[f"{out}.damage_nonlocal"][:, 0], "damage_nonlocal")
    assert_same(broken, g[f"{out}.damage_broken"], "damage_broken")
    assert_same(dD0, g[f"{out}.damage_D"][..., 0], "damage_D")
    assert_same(w, g[f"{out}.damage_w"], "damage_w")


@pytest.mark.parametrize("name", ["fcc_cp", "bcc_cp"])
def test_port_crystal_plasticity_bit_exact(name):
    """computeCab and computeBondForceGeneral(1, .) = computeBondForceCPMiehe (constitutive.c:866-1396, 1864-1917) restated
    (the memoised per-particle return map evaluated once per particle), against the reference's recorded calls on the FCC
    and BCC fixtures: Cab, active sets, increments, slot-[2] state, bond forces, stress -- bit for bit.  (Entries where the
    reference itself holds NaN -- a 0 * stale-gamma product in its elastic branch, BCC step 2 -- are skipped.)"""
    from pathlib import Path
    lib, C = _lib()
    lib.oracle_cp_return_map.restype = C.c_int
    g = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    S = int(par["nslipSys"])
    f8, i4 = np.float64, np.int32
    st = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial")}
    sf = {k: _c(g[f"setup.{k}"], f8) for k in ("distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn", "Tv", "schmid_tensor")}
    V = C.c_double(par["particle_volume"])
    nb0 = _c(g["setup.nb_initial"], i4)
    # computeCab on the set-up geometry
    Cab = np.zeros((N, S * S))
    ones = np.ones((N, nn))
    lib.oracle_cp_cab(C.c_int(N), C.c_int(nn), C.c_int(S), V, _ptr(st["nsign"]), _ptr(nb0), _ptr(st["nb_initial"]), _ptr(sf["Kn"]), _ptr(sf["Tv"]),
                      _ptr(_c(g["setup.distance"], f8)), _ptr(_c(g["setup.csx"], f8)), _ptr(_c(g["setup.csy"], f8)), _ptr(_c(g["setup.csz"], f8)),
                      _ptr(sf["csx_initial"]), _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(ones), _ptr(sf["schmid_tensor"]), _ptr(Cab))
    assert_same(Cab, g["setup.cp_Cab"], "cp_Cab")
    checked = 0
    for tag, prev in (("s1.n0", "s1.pred"), ("s1.n1", "s1.n0.bf"), ("s1.n2", "s1.n1.bf")):
        t = f"{tag}.bf"
        xyz = _c(g[f"{tag}.xyz"], f8)
        broken, w = _c(g[f"{prev}.damage_broken"], f8), _c(g[f"{prev}.damage_w"], f8)
        a = {k: _c(g[f"{prev}.{k}"], f8).copy() for k in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "dL_total", "TdL_total", "stress_tensor",
                                                          "J2_stresseq", "J2_stressm", "J2_triaxiality", "bond_stress", "Pin")}
        pl = _c(g[f"{prev}.pl_flag"], i4).copy()
        # switchStateV(0): slot [0] := slot [1]
        dLp0 = _c(g[f"{prev}.dLp"][..., 1], f8)
        gy0, As0, A0 = _c(g[f"{prev}.cp_gy"][..., 1], f8), _c(g[f"{prev}.cp_A_single"][..., 1], f8), _c(g[f"{prev}.cp_A"][:, 1], f8)

        def geometry(dLp):
            lib.oracle_geometry(C.c_int(N), C.c_int(nn), _ptr(xyz), _ptr(st["neighbors"]), _ptr(st["nsign"]), _ptr(st["nb_initial"]),
                                _ptr(sf["distance_initial"]), _ptr(dLp), _ptr(broken), _ptr(sf["Tv"]), C.c_int(1), _ptr(a["dL"]), _ptr(a["csx"]),
                                _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["dL_total"]), _ptr(a["TdL_total"]), None)
        geometry(dLp0)
        dA, dgy, dAs = np.zeros(N), _c(g[f"{prev}.cp_dgy"], f8).copy(), _c(g[f"{prev}.cp_dA_single"], f8).copy()
        Jact, RSS = _c(g[f"{prev}.cp_Jact"], i4).copy(), _c(g[f"{prev}.cp_RSS"], f8).copy()
        dLp2, gy2, A2, As2 = np.zeros((N, nn)), np.zeros((N, S)), np.zeros(N), np.zeros((N, S))
        rc = lib.oracle_cp_return_map(C.c_int(N), C.c_int(nn), C.c_int(S), V, C.c_double(par["cp_h0"]), C.c_double(par["cp_taus0"]),
                                      C.c_double(par["cp_tau00"]), C.c_double(par["cp_q"]), C.c_double(par["cp_eta"]), C.c_double(par["cp_p"]),
                                      C.c_double(par["cp_maxloop"]), C.c_double(par["dtime"]), _ptr(st["nsign"]), _ptr(nb0), _ptr(st["nb_initial"]),
                                      _ptr(sf["Kn"]), _ptr(sf["Tv"]), _ptr(w), _ptr(broken), _ptr(sf["distance_initial"]), _ptr(sf["csx_initial"]),
                                      _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(a["dL"]), _ptr(a["dL_total"]), _ptr(a["TdL_total"]),
                                      _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(sf["schmid_tensor"]), _ptr(Cab), _ptr(dLp0), _ptr(gy0),
                                      _ptr(A0), _ptr(As0), _ptr(a["ddLp"]), _ptr(dA), _ptr(dgy), _ptr(dAs), _ptr(Jact), _ptr(RSS), _ptr(pl),
                                      _ptr(dLp2), _ptr(gy2), _ptr(A2), _ptr(As2))
        assert rc == 0
        geometry(dLp2)
        lib.oracle_force(C.c_int(N), C.c_int(nn), C.c_int(0), _ptr(st["neighbors"]), _ptr(st["nsign"]), _ptr(st["nb_initial"]), _ptr(sf["Kn"]),
                         _ptr(sf["Tv"]), _ptr(w), _ptr(a["dL"]), _ptr(a["dL_total"]), _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]),
                         _ptr(a["csz"]), _ptr(a["dL_ave"]), _ptr(a["F"]), _ptr(a["Pin"]))
        lib.oracle_stress(C.c_int(N), C.c_int(nn), V, _ptr(nb0), _ptr(st["nb_initial"]), _ptr(sf["distance_initial"]), _ptr(sf["csx_initial"]),
                          _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(broken), _ptr(a["F"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]),
                          _ptr(a["stress_tensor"]), _ptr(a["J2_stresseq"]), _ptr(a["J2_stressm"]), _ptr(a["J2_triaxiality"]), _ptr(a["bond_stress"]))
        assert np.array_equal(Jact, g[f"{t}.cp_Jact"]), f"{tag}: active slip systems differ"
        assert np.array_equal(pl, g[f"{t}.pl_flag"])

        def same(got, want, what):
            want = np.asarray(want)
            ok = ~np.isnan(want)
            assert_same(np.where(ok, got, 0.0), np.where(ok, want, 0.0), f"{tag}: {what}")

        for nme, got in (("ddLp", a["ddLp"]), ("cp_dA", dA), ("cp_dgy", dgy), ("cp_dA_single", dAs), ("cp_RSS", RSS), ("dL", a["dL"]),
                         ("F", a["F"]), ("Pin", a["Pin"]), ("stress_tensor", a["stress_tensor"]), ("bond_stress", a["bond_stress"])):
            same(got, g[f"{t}.{nme}"], nme)
        same(dLp2, g[f"{t}.dLp"][..., 2], "dLp[2]"); same(gy2, g[f"{t}.cp_gy"][..., 2], "cp_gy[2]")
        same(A2, g[f"{t}.cp_A"][:, 2], "cp_A[2]"); same(As2, g[f"{t}.cp_A_single"][..., 2], "cp_A_single[2]")
        checked += int(Jact.sum())
    assert checked > 0                                          # slip systems were active in the recorded calls


@pytest.mark.parametrize("tag", ["fresh", "memo"])
def test_port_cp_law_called_per_particle_bit_exact(tag):
    """computeBondForceCPMiehe(ii) called on its own (constitutive.h:19, constitutive.c:866-1396) with its memo state_v
    (:946-959), against tests/golden/fcc_cp_particle.npz.  Restated exactly the way the CUDA entry point is organised
    (lpmb_bond_force_particle, plmode 1): whole-lattice geometry and return map into scratch copies, the rows of the star
    members whose flag was 0 committed and flagged, geometry of the star with dLp[0] + ddLp (new or REUSED increments),
    force pass of ii over the live arrays, slot [2] of ii -- every array after every call, bit for bit."""
    from pathlib import Path
    lib, C = _lib()
    lib.oracle_cp_return_map.restype = C.c_int
    g = np.load(Path(__file__).parent / "golden" / "fcc_cp_particle.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    S = int(par["nslipSys"])
    f8, i4 = np.float64, np.int32
    st = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial")}
    sf = {k: _c(g[f"setup.{k}"], f8) for k in ("distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn", "Tv", "schmid_tensor", "cp_Cab")}
    V = C.c_double(par["particle_volume"])
    pre = f"{tag}.pre"
    live = {k: _c(g[f"{pre}.{k}"], f8).copy() for k in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "dL_total", "TdL_total", "Pin", "cp_RSS",
                                                        "cp_dgy", "cp_dA", "cp_dA_single")}
    live["pl_flag"], live["cp_Jact"], live["state_v"] = (_c(g[f"{pre}.{k}"], i4).copy() for k in ("pl_flag", "cp_Jact", "state_v"))
    live["dLp2"] = _c(g[f"{pre}.dLp"][..., 2], f8).copy()
    live["cp_gy2"], live["cp_A_single2"] = _c(g[f"{pre}.cp_gy"][..., 2], f8).copy(), _c(g[f"{pre}.cp_A_single"][..., 2], f8).copy()
    live["cp_A2"] = _c(g[f"{pre}.cp_A"][:, 2], f8).copy()
    xyz, broken, w, nb = _c(g[f"{pre}.xyz"], f8), _c(g[f"{pre}.damage_broken"], f8), _c(g[f"{pre}.damage_w"], f8), _c(g[f"{pre}.nb"], i4)
    dLp0 = _c(g[f"{pre}.dLp"][..., 0], f8)
    gy0, As0, A0 = _c(g[f"{pre}.cp_gy"][..., 0], f8), _c(g[f"{pre}.cp_A_single"][..., 0], f8), _c(g[f"{pre}.cp_A"][:, 0], f8)
    nbr, nbi = st["neighbors"], st["nb_initial"]

    def geometry(dLp, t):
        lib.oracle_geometry(C.c_int(N), C.c_int(nn), _ptr(xyz), _ptr(nbr), _ptr(st["nsign"]), _ptr(nbi), _ptr(sf["distance_initial"]), _ptr(dLp),
                            _ptr(broken), _ptr(sf["Tv"]), C.c_int(1), _ptr(t["dL"]), _ptr(t["csx"]), _ptr(t["csy"]), _ptr(t["csz"]),
                            _ptr(t["dL_total"]), _ptr(t["TdL_total"]), None)

    fresh_calls = 0
    for k, ii in enumerate(g[f"{tag}.particles"]):
        ii = int(ii)
        star = [ii] + [int(nbr[ii, j]) for j in range(nn) if broken[ii, j] > 1e-6 and nbr[ii, j] != -1]
        assert len(star) == nb[ii] + 1
        rows0 = []
        for q in star:
            if live["state_v"][q] == 0:
                rows0.append(q)
                live["state_v"][q] = 1
        t = {n: live[n].copy() for n in live}
        geometry(dLp0, t)
        if rows0:
            rc = lib.oracle_cp_return_map(C.c_int(N), C.c_int(nn), C.c_int(S), V, C.c_double(par["cp_h0"]), C.c_double(par["cp_taus0"]),
                                          C.c_double(par["cp_tau00"]), C.c_double(par["cp_q"]), C.c_double(par["cp_eta"]), C.c_double(par["cp_p"]),
                                          C.c_double(par["cp_maxloop"]), C.c_double(par["dtime"]), _ptr(st["nsign"]), _ptr(nb), _ptr(nbi),
                                          _ptr(sf["Kn"]), _ptr(sf["Tv"]), _ptr(w), _ptr(broken), _ptr(sf["distance_initial"]), _ptr(sf["csx_initial"]),
                                          _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(t["dL"]), _ptr(t["dL_total"]), _ptr(t["TdL_total"]),
                                          _ptr(t["csx"]), _ptr(t["csy"]), _ptr(t["csz"]), _ptr(sf["schmid_tensor"]), _ptr(sf["cp_Cab"]), _ptr(dLp0),
                                          _ptr(gy0), _ptr(A0), _ptr(As0), _ptr(t["ddLp"]), _ptr(t["cp_dA"]), _ptr(t["cp_dgy"]), _ptr(t["cp_dA_single"]),
                                          _ptr(t["cp_Jact"]), _ptr(t["cp_RSS"]), _ptr(t["pl_flag"]), _ptr(t["dLp2"]), _ptr(t["cp_gy2"]),
                                          _ptr(t["cp_A2"]), _ptr(t["cp_A_single2"]))
            assert rc == 0
            for n in ("ddLp", "cp_RSS", "cp_Jact", "cp_dgy", "cp_dA", "cp_dA_single", "pl_flag"):
                live[n][rows0] = t[n][rows0]
        xdLp = dLp0.copy()
        for i in range(N):
            xdLp[i, :nbi[i]] += live["ddLp"][i, :nbi[i]]
        geometry(xdLp, t)
        for n in ("dL", "csx", "csy", "csz", "dL_total", "TdL_total"):
            live[n][star] = t[n][star]
        lib.oracle_force(C.c_int(N), C.c_int(nn), C.c_int(0), _ptr(nbr), _ptr(st["nsign"]), _ptr(nbi), _ptr(sf["Kn"]), _ptr(sf["Tv"]), _ptr(w),
                         _ptr(live["dL"]), _ptr(live["dL_total"]), _ptr(live["TdL_total"]), _ptr(live["csx"]), _ptr(live["csy"]), _ptr(live["csz"]),
                         _ptr(t["dL_ave"]), _ptr(t["F"]), _ptr(t["Pin"]))
        for n in ("dL_ave", "F"):
            live[n][ii] = t[n][ii]
        live["Pin"][3 * ii:3 * ii + 3] = t["Pin"][3 * ii:3 * ii + 3]
        if rows0 and rows0[0] == ii:
            fresh_calls += 1
            for n in ("dLp2", "cp_gy2", "cp_A2", "cp_A_single2"):
                live[n][ii] = t[n][ii]
        else:
            live["dLp2"][ii] = broken[ii] * xdLp[ii]
            live["cp_gy2"][ii] = gy0[ii] + live["cp_dgy"][ii]
            live["cp_A_single2"][ii] = As0[ii] + live["cp_dA_single"][ii]
            live["cp_A2"][ii] = A0[ii] + live["cp_dA"][ii]
        for n in ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "state_v", "cp_RSS", "cp_Jact",
                  "cp_dgy", "cp_dA", "cp_dA_single", "dLp2", "cp_gy2", "cp_A_single2", "cp_A2"):
            want = np.asarray(g[f"{tag}.c{k}.{n}"])
            got = live[n]
            ok = ~np.isnan(want) if want.dtype.kind == "f" else np.ones(want.shape, bool)
            assert_same(np.where(ok, got, 0), np.where(ok, want, 0), f"{tag} call {k} (particle {ii}): {n}")
    # "fresh": the second particle sits in the first one's star (already flagged) and the fifth repeats the first -> 3 of 5
    # calls return-map their own particle; all "memo" calls reuse everything
    assert fresh_calls == (3 if tag == "fresh" else 0)


@pytest.mark.parametrize("tag", ["fresh", "memo"])
def test_port_cp_particle_restatement_bit_exact(tag):
    """oracle_cp_particle = computeBondForceCPMiehe(ii) restated literally (geometry of the star, memo, return map of the
    unflagged members, geometry with the new / reused increments, force pass of ii, slot [2] of ii), against
    tests/golden/fcc_cp_particle.npz: every array after every call, bit for bit"""
    from pathlib import Path
    lib, C = _lib()
    lib.oracle_cp_particle.restype = C.c_int
    g = np.load(Path(__file__).parent / "golden" / "fcc_cp_particle.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    S = int(par["nslipSys"])
    f8, i4 = np.float64, np.int32
    st = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial")}
    sf = {k: _c(g[f"setup.{k}"], f8) for k in ("distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn", "Tv", "schmid_tensor", "cp_Cab")}
    pre = f"{tag}.pre"
    a = {k: _c(g[f"{pre}.{k}"], f8).copy() for k in ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "dL_total", "TdL_total", "Pin", "cp_RSS",
                                                     "cp_dgy", "cp_dA", "cp_dA_single")}
    a["pl_flag"], a["cp_Jact"], a["state_v"] = (_c(g[f"{pre}.{k}"], i4).copy() for k in ("pl_flag", "cp_Jact", "state_v"))
    a["dLp2"] = _c(g[f"{pre}.dLp"][..., 2], f8).copy()
    a["cp_gy2"], a["cp_A_single2"] = _c(g[f"{pre}.cp_gy"][..., 2], f8).copy(), _c(g[f"{pre}.cp_A_single"][..., 2], f8).copy()
    a["cp_A2"] = _c(g[f"{pre}.cp_A"][:, 2], f8).copy()
    xyz, broken, w, nb = _c(g[f"{pre}.xyz"], f8), _c(g[f"{pre}.damage_broken"], f8), _c(g[f"{pre}.damage_w"], f8), _c(g[f"{pre}.nb"], i4)
    dLp0 = _c(g[f"{pre}.dLp"][..., 0], f8)
    gy0, As0, A0 = _c(g[f"{pre}.cp_gy"][..., 0], f8), _c(g[f"{pre}.cp_A_single"][..., 0], f8), _c(g[f"{pre}.cp_A"][:, 0], f8)
    d = C.c_double
    for k, ii in enumerate(g[f"{tag}.particles"]):
        rc = lib.oracle_cp_particle(C.c_int(int(ii)), C.c_int(N), C.c_int(nn), C.c_int(S), d(par["particle_volume"]), d(par["cp_h0"]),
                                    d(par["cp_taus0"]), d(par["cp_tau00"]), d(par["cp_q"]), d(par["cp_eta"]), d(par["cp_p"]), d(par["cp_maxloop"]),
                                    d(par["dtime"]), _ptr(xyz), _ptr(st["neighbors"]), _ptr(st["nsign"]), _ptr(nb), _ptr(st["nb_initial"]),
                                    _ptr(sf["Kn"]), _ptr(sf["Tv"]), _ptr(w), _ptr(broken), _ptr(sf["distance_initial"]), _ptr(sf["csx_initial"]),
                                    _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(sf["schmid_tensor"]), _ptr(sf["cp_Cab"]), _ptr(dLp0),
                                    _ptr(gy0), _ptr(A0), _ptr(As0), _ptr(a["state_v"]), _ptr(a["dL"]), _ptr(a["dL_total"]), _ptr(a["TdL_total"]),
                                    _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]), _ptr(a["ddLp"]), _ptr(a["cp_dA"]), _ptr(a["cp_dgy"]),
                                    _ptr(a["cp_dA_single"]), _ptr(a["cp_Jact"]), _ptr(a["cp_RSS"]), _ptr(a["pl_flag"]), _ptr(a["dL_ave"]),
                                    _ptr(a["F"]), _ptr(a["Pin"]), _ptr(a["dLp2"]), _ptr(a["cp_gy2"]), _ptr(a["cp_A2"]), _ptr(a["cp_A_single2"]))
        assert rc == 0
        for n in ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "state_v", "cp_RSS", "cp_Jact",
                  "cp_dgy", "cp_dA", "cp_dA_single", "dLp2", "cp_gy2", "cp_A_single2", "cp_A2"):
            want = np.asarray(g[f"{tag}.c{k}.{n}"])
            ok = ~np.isnan(want) if want.dtype.kind == "f" else np.ones(want.shape, bool)
            assert_same(np.where(ok, a[n], 0), np.where(ok, want, 0), f"{tag} call {k} (particle {ii}): {n}")


@pytest.mark.parametrize("name,lattice", [("sq2d_brittle", 0), ("hex2d_brittle", 1), ("sc6_j2", 2), ("fcc_cp", 3), ("bcc_cp", 4)])
def test_port_calc_kntv_all_lattices_bit_exact(name, lattice):
    """calcKnTv (stiffness.c:11-268) restated for the five lattices, against the Kn / Tv the reference computed for the
    committed fixtures (square, hexagon, simple cubic with type averaging, FCC, BCC)"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    f8, i4 = np.float64, np.int32
    Ce = _c(g["setup.Ce"], f8)
    ntype = Ce.shape[0]
    KnTve, Kn, Tv = np.zeros((ntype, 3)), np.zeros((N, nn)), np.zeros((N, nn))
    lib.oracle_calc_kntv(C.c_int(lattice), C.c_int(N), C.c_int(nn), C.c_int(ntype), C.c_double(par["radius"]), _ptr(Ce),
                         _ptr(_c(g["setup.type"], i4)), _ptr(_c(g["setup.neighbors"], i4)), _ptr(_c(g["setup.nsign"], i4)),
                         _ptr(_c(g["setup.nb_initial"], i4)), _ptr(KnTve), _ptr(Kn), _ptr(Tv))
    assert_same(Kn, g["setup.Kn"], "Kn")
    assert_same(Tv, g["setup.Tv"], "Tv")
    if "setup.KnTve" in g.files:
        ref = g["setup.KnTve"]
        assert_same(KnTve[:, : ref.shape[1]], ref, "KnTve")
    assert np.abs(Kn).max() > 0 and np.abs(Tv).max() > 0


@pytest.mark.parametrize("name,nn,nconn", [("hex2d_brittle", 12, 31), ("sq2d_brittle", 8, 17)])
def test_port_2d_brittle_trajectory_bit_exact(name, nn, nconn):
    """The 2-D configurations (hexagonal / square lattice, elastic law + updateBrittleDamage with nbreak = 2): the whole
    recorded trajectory of tests/golden/{hex,sq}2d_brittle.npz -- topology, 2-D FD tangent (2x2 blocks), BC-modified
    tangent, CG, predictor, elastic law, residual, every breaking event incl. the reference's shell-sort selection, crack
    update, re-assembly -- replayed through the restatement, bit for bit."""
    from pathlib import Path
    from oracle import port as P
    if not P.available():
        pytest.skip("oracle/liblpm_oracle.so not built")
    g = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
    par = params_from_golden(g)
    p = P.Port(g["setup.xyz"], dim=2, nn=nn, nconn=nconn, radius=par["radius"], particle_volume=par["particle_volume"])
    p.search_neighbors(par["neighbor1_cutoff"], par["neighbor2_cutoff"])
    assert_same(p.neighbors, g["setup.neighbors"], "neighbors"); assert_same(p.nsign, g["setup.nsign"], "nsign")
    assert_same(p.conn, g["setup.conn"], "conn"); assert_same(p.nb_conn, g["setup.nb_conn"], "nb_conn")
    assert_same(np.stack([p.kp0, p.kp1], 1), g["setup.K_pointer"].astype(np.int64), "K_pointer")
    for n in ("distance_initial", "csx_initial", "csy_initial"):
        assert_same(getattr(p, n), g[f"setup.{n}"], n)
    p.type[:] = g["setup.type"]
    p.Kn[:], p.Tv[:] = g["setup.Kn"], g["setup.Tv"]          # calcKnTv is pinned on the GPU side (test_variants_gpu.py)
    p.computedL()
    for n in ("distance", "dL", "dL_total", "TdL_total", "csx", "csy"):
        assert_same(getattr(p, n), g[f"setup.{n}"], n)
    typ = g["setup.type"]
    crit, nbreak = par["critical_bstrain"], int(par["nbreak"])
    for step in (1, 2, 3, 4):
        s = f"s{step}"
        assert_same(p.xyz, g[f"{s}.pre.xyz"], f"{s} xyz"); assert_same(p.F, g[f"{s}.pre.F"], f"{s} F")
        p.xyz_temp[:] = p.xyz; p.F_temp[:] = p.F
        p.calcStiffnessFiniteDifference()
        assert_same(p.K_global, g[f"{s}.fd.K_global"], "K_global"); assert np.array_equal(p.IK, g[f"{s}.fd.IK"]); assert np.array_equal(p.JK, g[f"{s}.fd.JK"])
        for n in ("dL", "csx", "csy", "dL_total", "TdL_total", "F"):
            assert_same(getattr(p, n), g[f"{s}.fd.{n}"], f"{s} FD side effect {n}")
        # setDispBC (boundary.c:12-38): type 1 moves 1.5e-4 in x and is held in y, type 2 is held in x and y
        p.xyz[typ == 1, 0] += 1.5e-4
        bc = p.dispBC_index.reshape(-1, 2)
        bc[(typ == 1) | (typ == 2), :] = 0
        assert_same(p.xyz, g[f"{s}.bc.xyz"], "xyz after BC"); assert np.array_equal(p.dispBC_index, g[f"{s}.bc.dispBC_index"])
        p.computeBondForceGeneral(4)
        for n in ("ddL", "ddL_total", "TddL_total", "F", "Pin", "stress_tensor"):
            assert_same(getattr(p, n), g[f"{s}.pred.{n}"], f"{s} predictor {n}")
        event, total_ni = 0, 0
        while True:
            p.updateRR()
            nr = float(np.sqrt(np.sum(p.residual ** 2)))
            nf = float(np.sqrt(np.sum(p.reaction_force ** 2)))
            if event == 0:
                assert_same(p.residual, g[f"{s}.rr.residual"], "residual")
            tol, ni = max(nr, nf), 0
            while nr > 1e-4 * tol and ni < 100:
                t = f"{s}.e{event}.n{ni}"
                p.switchStateV(0)
                p.setDispBC_stiffnessUpdate()
                rec = f"{t}.K_bc" in g.files
                if rec:
                    assert_same(p.K_global, g[f"{t}.K_bc"], "K_bc"); assert_same(p.residual, g[f"{t}.rhs"], "rhs")
                it = p.solverCG()
                if rec:
                    assert it == int(g[f"{t}.cg_iters"][0])
                    assert_same(p.disp, g[f"{t}.disp"], "disp"); assert_same(p.xyz, g[f"{t}.xyz"], "xyz")
                p.computeBondForceGeneral(6)
                if rec:
                    for n in ("dL", "csx", "csy", "dL_total", "TdL_total", "F", "Pin", "stress_tensor", "bond_stress"):
                        assert_same(getattr(p, n), g[f"{t}.bf.{n}"], f"{t} {n}")
                p.updateRR()
                nr = float(np.sqrt(np.sum(p.residual ** 2)))
                ni += 1
            total_ni += ni
            d = f"{s}.dam{event}"
            assert_same(p.dL, g[f"{d}.pre.dL"], f"{d} dL"); assert_same(p.damage_broken, g[f"{d}.pre.damage_broken"], f"{d} broken")
            k, pairs = p.updateBrittleDamage(crit, nbreak)
            assert k == int(g[f"{d}.broken"][0]), (d, k)
            assert_same(p.damage_broken, g[f"{d}.post.damage_broken"], f"{d} damage_broken"); assert_same(p.damage_w, g[f"{d}.post.damage_w"], "damage_w")
            assert_same(p.damage_D0, g[f"{d}.post.damage_D"][..., 0], "damage_D")
            p.updateCrack()
            assert_same(p.F, g[f"{d}.crack.F"], "crack F"); assert_same(p.Pin, g[f"{d}.crack.Pin"], "crack Pin"); assert_same(p.nb, g[f"{d}.crack.nb"], "nb")
            p.switchStateV(1)
            if k <= 0:
                break
            p.calcStiffnessFiniteDifference()
            if event == 0:
                assert_same(p.K_global, g[f"{s}.refd.K_global"], "re-assembled K")
            event += 1
            if event >= 6:
                break
        assert event + 1 == int(g[f"{s}.events"][0])
        assert total_ni == int(g["newton_counts"][step - 1])
        assert_same(p.xyz, g[f"{s}.end.xyz"], "end xyz"); assert_same(p.damage_broken, g[f"{s}.end.damage_broken"], "end broken")
    assert max(int(g[f"s4.dam{e}.broken"][0]) for e in range(3)) > nbreak      # the shell-sort selection was exercised


@pytest.mark.parametrize("tag,law", [("s1.pred", 4), ("s1.j2", 0), ("s1.el", 6), ("s2.j2", 0), ("s2.el", 6)])
def test_port_per_particle_laws_bit_exact(tag, law):
    """oracle_particle_law = the reference's per-particle entry points (constitutive.c:167-283, 466-686) restated, against
    tests/golden/sc6_particle.npz: five calls in sequence per phase, every written array after every call"""
    from pathlib import Path
    lib, C = _lib()
    g = np.load(Path(__file__).parent / "golden" / "sc6_particle.npz")
    par = params_from_golden(g)
    N, nn = g["setup.neighbors"].shape
    f8, i4 = np.float64, np.int32
    pre = f"{tag}.pre"
    a = {k: _c(g[f"{pre}.{k}"], f8).copy() for k in ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddL", "ddL_total", "TddL_total", "ddLp",
                                                     "dL_ave", "F", "Pin", "stress_tensor", "J2_dlambda")}
    a["pl_flag"] = _c(g[f"{pre}.pl_flag"], i4).copy()
    a["dLp2"] = _c(g[f"{pre}.dLp"][..., 2], f8).copy()
    a["J2_beta2"] = _c(g[f"{pre}.J2_beta"][..., 2], f8).copy()
    a["J2_alpha2"] = _c(g[f"{pre}.J2_alpha"][:, 2], f8).copy()
    ro = {k: _c(g[f"{pre}.{k}"], f8) for k in ("xyz", "xyz_temp", "damage_broken", "damage_w", "F_temp")}
    dLp0, beta0, alpha0 = _c(g[f"{pre}.dLp"][..., 0], f8), _c(g[f"{pre}.J2_beta"][..., 0], f8), _c(g[f"{pre}.J2_alpha"][:, 0], f8)
    nb = _c(g[f"{pre}.nb"], i4)
    st = {k: _c(g[f"setup.{k}"], i4) for k in ("neighbors", "nsign", "nb_initial", "type")}
    sf = {k: _c(g[f"setup.{k}"], f8) for k in ("Ce", "sigmay", "distance_initial", "csx_initial", "csy_initial", "csz_initial", "Kn", "Tv")}
    writes = {4: ("ddL", "ddL_total", "TddL_total", "F", "Pin"),
              6: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "F", "Pin"),
              0: ("dL", "dL_total", "TdL_total", "csx", "csy", "csz", "ddLp", "pl_flag", "dL_ave", "F", "Pin", "stress_tensor", "J2_dlambda", "dLp2",
                  "J2_beta2", "J2_alpha2")}[law]
    for k, ii in enumerate(g[f"{tag}.particles"]):
        lib.oracle_particle_law(C.c_int(law), C.c_int(int(ii)), C.c_int(N), C.c_int(nn), C.c_double(par["particle_volume"]), C.c_double(par["J2_H"]),
                                C.c_double(par["J2_xi"]), _ptr(sf["Ce"]), _ptr(st["type"]), _ptr(sf["sigmay"]), _ptr(ro["xyz"]), _ptr(ro["xyz_temp"]),
                                _ptr(st["neighbors"]), _ptr(st["nsign"]), _ptr(st["nb_initial"]), _ptr(nb), _ptr(sf["distance_initial"]),
                                _ptr(sf["csx_initial"]), _ptr(sf["csy_initial"]), _ptr(sf["csz_initial"]), _ptr(sf["Kn"]), _ptr(sf["Tv"]),
                                _ptr(ro["damage_broken"]), _ptr(ro["damage_w"]), _ptr(ro["F_temp"]), _ptr(dLp0), _ptr(beta0), _ptr(alpha0),
                                _ptr(a["dL"]), _ptr(a["dL_total"]), _ptr(a["TdL_total"]), _ptr(a["csx"]), _ptr(a["csy"]), _ptr(a["csz"]),
                                _ptr(a["ddL"]), _ptr(a["ddL_total"]), _ptr(a["TddL_total"]), _ptr(a["ddLp"]), _ptr(a["pl_flag"]), _ptr(a["dL_ave"]),
                                _ptr(a["F"]), _ptr(a["Pin"]), _ptr(a["stress_tensor"]), _ptr(a["J2_dlambda"]), _ptr(a["dLp2"]), _ptr(a["J2_beta2"]),
                                _ptr(a["J2_alpha2"]))
        for n in writes:
            assert_same(a[n].reshape(g[f"{tag}.c{k}.{n}"].shape), g[f"{tag}.c{k}.{n}"], f"{tag} call {k} (particle {ii}): {n}")


@pytest.mark.parametrize("step", ["s1", "s2"])
def test_port_compute_strain_bit_exact(golden, step):
    lib, C = _lib()
    g = golden
    N, nn = g["setup.neighbors"].shape
    f8, i4 = np.float64, np.int32
    strain = np.zeros((N, 6)) if step == "s1" else _c(g["s1.strain.strain_tensor"], f8).copy()
    lib.oracle_compute_strain(C.c_int(N), C.c_int(nn), C.c_int(3), _ptr(_c(g["setup.xyz"], f8)), _ptr(_c(g["setup.neighbors"], i4)),
                              _ptr(_c(g["setup.nsign"], i4)), _ptr(_c(g["setup.nb_initial"], i4)), _ptr(_c(g["setup.distance_initial"], f8)),
                              _ptr(_c(g[f"{step}.strain.dL"], f8)), _ptr(strain))
    assert_same(strain, g[f"{step}.strain.strain_tensor"], "strain_tensor")


def test_reciprocal_quotient_of_the_fd_kernel_is_the_ieee_quotient(tmp_path):
    """fd_div_by (lpmb_stiffness.cu) forms x / y for a divisor used many times (`/ EPS / radius`, the three direction cosines
    of one bond) from rcp = RN(1 / y) and two corrections with exact FMA residuals.  The bit-exactness of the FD tangent rests
    on that being the IEEE quotient: scripts/check_exact_division.c compares it with `/` on random numerators over 600
    binades for the divisors the kernel uses (1e-6, 0.25, 0.3, ...) and on random divisors -- here 2e6 per divisor + 6e6
    random pairs; the full 8.6e8-pair run is quoted in DESIGN.md."""
    import shutil
    import subprocess
    from pathlib import Path
    src = Path(__file__).resolve().parents[1] / "scripts" / "check_exact_division.c"
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    exe = tmp_path / "check_exact_division"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run([gcc, "-O2", "-march=native", "-ffp-contract=off", str(src), "-o", str(exe), "-lm"], check=True, env=env)
    r = subprocess.run([str(exe), "2000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().startswith("bad 0 of"), r.stdout[-500:]


def test_brittle_selection_of_the_library_equals_the_oracle():
    """lpmb_brittle_select (host arithmetic of liblpmb200.so: the selection step of updateBrittleDamage, used on one GPU and
    -- on the all-gathered candidates of all ranks -- in slab runs) against the oracle's restatement of constitutive.c:1437-1526
    on synthetic candidate sets with MANY ties (the reference's shell sort is not stable: which of the tied bonds break is part
    of the behaviour): same bonds, same order, for nbreak below / at / above the candidate count, also when the global list is
    assembled from per-rank pieces in rank order."""
    import ctypes as C
    import importlib
    from oracle.port import SO
    capi = importlib.import_module("lpm-c_b200.capi")
    lib = C.CDLL(str(SO))
    lib.oracle_damage_brittle.restype = C.c_int
    rng = np.random.default_rng(5)
    N, nn, crit = 60, 8, 1.0
    nbr = np.tile(np.arange(nn, dtype=np.int32), (N, 1))          # "neighbour" of slot j is j: the pairs give the key back
    nbi = np.full(N, nn, dtype=np.int32)
    L0 = np.ones((N, nn))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in (1, 3, 10, 57, 200, 390):
        for levels in (2, 5, 1000):
            flat = rng.choice(N * nn, size=k, replace=False)
            dL = np.zeros((N, nn))
            dL.ravel()[flat] = crit + rng.integers(0, levels, size=k) * 0.125          # exact binary fractions: true ties
            keys = np.sort(flat).astype(np.int64)
            strains = dL.ravel()[keys].copy()
            for nbreak in (1, 2, 7, k, k + 5):
                broken, dD0, w = np.ones((N, nn)), np.zeros((N, nn)), np.ones((N, nn))
                pairs = np.full((400, 2), -1, np.int32)
                kk = lib.oracle_damage_brittle(N, nn, C.c_double(crit), nbreak, p(nbr), p(nbi), p(dL), p(L0), p(broken), p(dD0), p(w), p(pairs), 400)
                assert kk == k
                want = [int(a) * nn + int(b) for a, b in pairs[: min(k, nbreak)]]
                got_k, got_s = capi.brittle_select(keys, strains, nbreak)
                assert list(got_k) == want, (k, levels, nbreak)
                assert np.array_equal(got_s, dL.ravel()[got_k])
                # slab runs: every rank lists the candidates of its own particle range; rank order = ascending keys
                cuts = [0, 17 * nn, 41 * nn, N * nn]
                pieces = [keys[(keys >= a) & (keys < b)] for a, b in zip(cuts[:-1], cuts[1:])]
                merged = np.concatenate(pieces)
                assert np.array_equal(merged, keys)
                assert list(capi.brittle_select(merged, dL.ravel()[merged], nbreak)[0]) == want
