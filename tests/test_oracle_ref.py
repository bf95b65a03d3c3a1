"""CPU: pin the oracle.

The reference has no tests or golden vectors; what exists are the known answers
the survey measured on the default configuration and printed by the reference's own code.  The
oracle build (unmodified reference sources + open MKL stand-in) must reproduce them, and the
committed golden vectors must be what that build produces.
"""
import numpy as np
import pytest


def test_default_case_known_answers(ref_c1):
    r = ref_c1["ref"]
    assert r.N == 9261                                  # "Particle number is 9261"
    kp = r.get("K_pointer")
    assert int(kp[-1, 1]) == 2203713                    # stiffness matrix size (lpmc_project.c:166)
    assert int(r.get("nb_initial").sum()) == 153720
    assert int(r.get("nb_conn").sum()) == 486627
    assert ref_c1["norm_residual"] == pytest.approx(2000.0 / np.sqrt(441.0), rel=1e-12)   # 95.2381
    assert ref_c1["norm_reaction"] == 0.0
    # lattice origin quirk
    x0 = r.get("xyz_initial")[0]
    assert x0[0] == -0.2 and abs(x0[1] + 0.0133288) < 1e-6
    # SC spring constants for type 0: KnTve = (14038.46, 14038.46, 389.957)
    np.testing.assert_allclose(r.get("KnTve")[0], [14038.46, 14038.46, 389.957], rtol=1e-6)


def test_default_case_cg_iteration_counts(ref_c1):
    """step-1 CG iteration counts 80 then 106; K is symmetric by construction and the
    unconstrained tangent annihilates rigid translations"""
    r = ref_c1["ref"]
    L = r.lib
    K, IK, JK = r.get("K_global"), r.get("IK"), r.get("JK")
    import scipy.sparse as sp
    n = 3 * r.N
    U = sp.csr_matrix((K, JK - 1, IK - 1), shape=(n, n))
    A = U + sp.triu(U, 1).T
    t = np.zeros(n)
    t[2::3] = 1.0
    assert np.abs(A @ t).max() < 1e-6 * np.abs(K).max()
    r.newton_iteration()
    assert L.lpmb_shim_last_itercount() == 80
    nr = r.newton_iteration()
    assert L.lpmb_shim_last_itercount() == 106
    assert nr < 1e-4 * ref_c1["norm_residual"]
    # mean displacement of the loaded layer after step 1: -1.27857453e-03 (result_disp.txt)
    typ = r.get("type")
    xyz, xyz0 = r.get("xyz"), r.get("xyz_initial")
    # the default driver reports type 2 (bottom layer is 2? no: type 1 = top, 2 = bottom; dtype=2)
    assert np.isfinite(xyz).all()


def test_threaded_cpu_baseline_solves_like_the_serial_reference(ref_c1):
    """bench.py times the reference with all host threads; the shim then runs a threaded symmetric SpMV on the stored
    triangle (chunked rows + overflow buffers).  Same matrix, same rhs: the threaded solve must agree with the serial one
    up to summation order -- same iteration count +-1, disp to 1e-9 -- also when the chunks are shorter than the band
    (12 threads on 27 783 rows with a band of ~2 700) and with fewer rows than threads' worth of work at the edges."""
    import scipy.sparse as sp
    r = ref_c1["ref"]
    L = r.lib
    L.switchStateV(0)
    L.setDispBC_stiffnessUpdate3D()
    xyz0, rhs = r.get("xyz"), r.get("residual")
    K, IK, JK = r.get("K_global"), r.get("IK"), r.get("JK")
    n = 3 * r.N
    U = sp.csr_matrix((K, JK - 1, IK - 1), shape=(n, n))
    A = U + sp.triu(U, 1).T
    L.solverCG()
    it1, d1 = L.lpmb_shim_last_itercount(), r.get("disp")
    assert it1 == 80
    assert np.linalg.norm(A @ d1 - rhs) <= 2e-4 * np.linalg.norm(rhs)
    try:
        for nt in (3, 12, 64):
            r.put("xyz", xyz0)
            r.put("residual", rhs)
            r.threads(nt)
            L.solverCG()
            it, d = L.lpmb_shim_last_itercount(), r.get("disp")
            assert abs(it - it1) <= 1, (nt, it, it1)
            assert np.linalg.norm(d - d1) <= 1e-9 * np.linalg.norm(d1), nt
    finally:
        r.threads(1)


def test_point_preconditioners_do_not_pay_on_the_lattice_tangent(ref_c1):
    """north_star speaks of a preconditioned CG; the reference's solverCG is unpreconditioned (solver.c:219-220, ipar[10] = 0)
    and the GPU solver replays it iteration for iteration.  This measures what the cheap (HBM-neutral) preconditioners would
    buy on the reference's own tangent: the diagonal blocks of a uniform lattice are all alike, so Jacobi / 3x3 block-Jacobi
    change the iteration count by a few percent only (80 -> 77 / 76 on the default case) while costing a vector pass per
    iteration and the iteration-count parity -- DESIGN.md section 3."""
    import scipy.sparse as sp
    r = ref_c1["ref"]
    L = r.lib
    L.switchStateV(0)
    L.setDispBC_stiffnessUpdate3D()
    K, IK, JK, rhs = r.get("K_global"), r.get("IK"), r.get("JK"), r.get("residual")
    n = 3 * r.N
    U = sp.csr_matrix((K, JK - 1, IK - 1), shape=(n, n))
    A = (U + sp.triu(U, 1).T).tocsr()

    def cg(M):
        x, res = np.zeros(n), rhs.copy()
        z = M(res)
        p, rz = z.copy(), res @ z
        thr = 1e-8 * (res @ res) + 1e-12
        for it in range(1, 1000):
            Ap = A @ p
            a = rz / (p @ Ap)
            x += a * p
            res -= a * Ap
            if res @ res <= thr:
                return it
            z = M(res)
            rz, rz_old = res @ z, rz
            p = z + (rz / rz_old) * p
        return 1000

    d = A.diagonal()
    blocks = np.stack([A[3 * i:3 * i + 3, 3 * i:3 * i + 3].toarray() for i in range(0, r.N, 1)])
    binv = np.linalg.inv(blocks)
    plain = cg(lambda v: v)
    jacobi = cg(lambda v: v / d)
    bjacobi = cg(lambda v: np.einsum("nij,nj->ni", binv, v.reshape(-1, 3)).ravel())
    assert plain == 80
    assert 0.9 * plain <= jacobi <= plain and 0.9 * plain <= bjacobi <= plain, (plain, jacobi, bjacobi)


@pytest.mark.parametrize("tag", ["C2", "C3", "C4", "C5src"])
def test_topology_known_answers_at_the_real_config_sizes(ref, tag):
    """particle count, sum of bonds, sum of conn blocks and nnz_upper of BASELINE configs 2-4
    at their REAL sizes (the committed fixtures of these lattices are small blocks), produced by the reference's own set-up
    code -- and the oracle port's restatement of neighbor.c:9-141 reproduces the lists bit for bit at that size."""
    from oracle import port as P
    r = ref
    if tag == "C2":     # examples/shear_hex_brittle.c:61,70,97-99,121-122
        r.setup_2d(lattice=1, box=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0), radius=3.2e-3, crack=(-0.5, 0.5, 0.5))
        want, dim, nn, nconn = (28170, 334370, 858256, 1744682), 2, 12, 31
    elif tag == "C3":   # examples/3_point_bending_sq_brittle.c:61,70,97-99,121-122 (notch of half-width 1.2 r)
        r.setup_2d(lattice=0, box=(0.0, 0.2, 0.0, 1.0, 0.0, 1.0), radius=2e-3, crack=(-0.5, 0.08, 0.5002), crack_w=1.2 * 2e-3,
                   critical_bstrain=2.7e-4)
        want, dim, nn, nconn = (12460, 97768, 206096, 424652), 2, 8, 17
    elif tag == "C5src":  # examples/CT_sc_ductile_nonlocal.c as shipped: carved, pre-cracked compact-tension specimen
        r.threads(8)      # its searches are OpenMP loops over disjoint rows (neighbor.c:13,48); the lists do not depend on it
        try:
            r.setup_ct_geometry()
        finally:
            r.threads(1)
        want, dim, nn, nconn = (75030, 1267032, 4072398, 18438336), 3, 18, 61
    else:               # examples/FCC_Al_R0.3_001_tension.c geometry through the default driver (lattice 3, r = 0.3, box 0..10)
        r.setup_fcc()
        want, dim, nn, nconn = (6912, 114192, 365016, 1652940), 3, 18, 61
    N = r.N
    nbr, nsign, conn = r.get("neighbors"), r.get("nsign"), r.get("conn")
    kp = r.get("K_pointer")
    got = (N, int(r.get("nb_initial").sum()), int(r.get("nb_conn").sum()), int(kp[N, 1]))
    assert got == want, got
    assert nbr.shape[1] == nn and conn.shape[1] == nconn
    if not P.available():
        pytest.skip("oracle/liblpm_oracle.so not built")
    p = P.Port(r.get("xyz_initial"), dim=dim, nn=nn, nconn=nconn, radius=r.gd("radius"), particle_volume=r.gd("particle_volume"))
    p.search_neighbors(r.gd("neighbor1_cutoff"), r.gd("neighbor2_cutoff"))
    assert np.array_equal(p.neighbors, nbr) and np.array_equal(p.nsign, nsign), "neighbour lists differ"
    assert np.array_equal(p.nb_initial, r.get("nb_initial"))
    assert np.array_equal(p.conn, conn) and np.array_equal(p.nb_conn, r.get("nb_conn")), "conn differs"
    assert np.array_equal(p.kp0[:N], kp[:N, 0]) and np.array_equal(p.kp1, kp[:, 1]), "K_pointer differs"
    assert np.array_equal(p.distance_initial, r.get("distance_initial"))


def test_golden_matches_reference_build(golden):
    """the committed fixture is bit-identical to what the oracle build produces today"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    # regenerate in a subprocess (the reference keeps global state; the session fixture may hold C1)
    import subprocess, sys, tempfile, shutil
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    with tempfile.TemporaryDirectory() as td:
        tmp = Path(td) / "golden"
        tmp.mkdir()
        shutil.copy(root / "tests" / "golden" / "make_golden.py", tmp / "make_golden.py")
        # make_golden resolves the repo root two levels up: emulate the layout
        (Path(td) / "tests").mkdir()
        shutil.move(str(tmp), str(Path(td) / "tests" / "golden"))
        for d in ("oracle",):
            (Path(td) / d).symlink_to(root / d)
        subprocess.run([sys.executable, str(Path(td) / "tests" / "golden" / "make_golden.py")], check=True,
                       stdout=subprocess.DEVNULL)
        new = np.load(Path(td) / "tests" / "golden" / "sc6_j2.npz")
        assert sorted(new.files) == sorted(golden.files)
        for k in golden.files:
            assert np.array_equal(new[k], golden[k]), k


def test_j2_energy_golden_matches_reference_build(tmp_path):
    """tests/golden/sc6_j2energy.npz (plmode 3) regenerates bit-identically from oracle/_ref"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    out = tmp_path / "j2e.npz"
    subprocess.run([sys.executable, str(gold / "make_golden_j2e.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT=str(out)))
    new, old = np.load(out), np.load(gold / "sc6_j2energy.npz")
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k
    # the law is really exercised: bisection results strictly inside (0, 1), reversed load indicator in step 2
    dl = old["s1.n0.bf.J2_dlambda"]
    assert (dl > 0).all() and dl.max() < 1.0
    assert (old["s2.n1.bf.J2_dlambda"] > 0).sum() > 50


def test_j2_iso_golden_matches_reference_build(tmp_path):
    """tests/golden/sc6_j2iso.npz (plmode 5 + local bond-wise damage) regenerates bit-identically"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    out = tmp_path / "j2iso.npz"
    subprocess.run([sys.executable, str(gold / "make_golden_j2iso.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT=str(out)))
    new, old = np.load(out), np.load(gold / "sc6_j2iso.npz")
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k
    # the law is active inside the call although its state is wiped afterwards: the elastic stretch it leaves in dL
    # differs from the purely geometric one; 18 bonds break in step 2
    x, nbr, L0 = old["s1.n0.bf.xyz"], old["setup.neighbors"], old["setup.distance_initial"]
    geo = np.linalg.norm(x[:, None, :] - x[np.maximum(nbr, 0)], axis=2) - L0
    assert np.abs(np.where(nbr >= 0, old["s1.n0.bf.dL"] - geo, 0.0)).max() > 1e-3
    assert np.abs(old["s1.n0.bf.dLp"]).max() == 0.0           # switchStateV(2) restored the never-written slot [2]
    assert int(old["s2.dam.broken"][0]) == 18


def test_damage_variants_golden_matches_reference_build(tmp_path):
    """tests/golden/sc6_damage_variants.npz (the two ductile-damage laws the dispatcher keeps commented out, SURVEY row
    a16) regenerates bit-identically; the recorded calls really break particles / bonds and freeze damaged particles"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    out = tmp_path / "dv.npz"
    subprocess.run([sys.executable, str(gold / "make_golden_damage_variants.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT=str(out)))
    new, old = np.load(out), np.load(gold / "sc6_damage_variants.npz")
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k
    assert [int(old[f"pwl.s{s}.broken"][0]) for s in (1, 2, 3)] == [0, 9, 16]
    assert [int(old[f"bwn.s{s}.broken"][0]) for s in (1, 2, 3)] == [0, 6, 264]
    thr = 0.02
    assert (old["bwn.s3.pre.damage_nonlocal"][:, 0] > thr).any()        # frozen branch (constitutive.c:1712-1717) taken


def test_per_particle_golden_matches_reference_build(tmp_path):
    """tests/golden/sc6_particle.npz (the per-particle law entry points called outside the dispatcher) regenerates
    bit-identically; every call changes rows of the star only"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    out = tmp_path / "pp.npz"
    subprocess.run([sys.executable, str(gold / "make_golden_particle.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT=str(out)))
    new, old = np.load(out), np.load(gold / "sc6_particle.npz")
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k
    nbr, broken = old["setup.neighbors"], old["s2.el.pre.damage_broken"]
    prev = old["s2.el.pre.dL"]
    for k, ii in enumerate(old["s2.el.particles"]):
        cur = old[f"s2.el.c{k}.dL"]
        star = {int(ii)} | {int(j) for j, b in zip(nbr[ii], broken[ii]) if j >= 0 and b > 1e-6}
        touched = set(np.flatnonzero((cur != prev).any(axis=1)).tolist())
        assert touched and touched <= star, (ii, touched - star)
        prev = cur


def test_2d_goldens_match_reference_build(tmp_path):
    """tests/golden/hex2d_brittle.npz / sq2d_brittle.npz (BASELINE configs 2 and 3 on a small box) regenerate bit-identically;
    the recorded runs break bonds and exercise the shell-sort selection (more candidates than nbreak)"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    subprocess.run([sys.executable, str(gold / "make_golden_2d.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT_DIR=str(tmp_path)))
    for name, nn, nconn in (("hex2d_brittle", 12, 31), ("sq2d_brittle", 8, 17)):
        new, old = np.load(tmp_path / f"{name}.npz"), np.load(gold / f"{name}.npz")
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            assert np.array_equal(new[k], old[k]), (name, k)
        assert old["setup.neighbors"].shape[1] == nn and old["setup.conn"].shape[1] == nconn
        cands = [int(old[k][0]) for k in old.files if k.endswith(".broken")]
        assert max(cands) > 2 and (old["s4.end.damage_broken"] == 0).sum() > 0


def test_bcc_cp_golden_matches_reference_build(tmp_path):
    """tests/golden/bcc_cp.npz (crystal plasticity on the BCC lattice: 14 neighbours, 41 conn, 24 slip systems) regenerates
    bit-identically; slip systems are active in both load steps"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import os, subprocess, sys
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    out = tmp_path / "bcc.npz"
    subprocess.run([sys.executable, str(gold / "make_golden_cp.py")], check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, LPMB_GOLDEN_OUT=str(out), LPMB_CP_LATTICE="4"))
    new, old = np.load(out), np.load(gold / "bcc_cp.npz")
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        # (the reference's law leaves NaN in cp_dA_single / cp_A_single of one intermediate Newton iteration of step 2 --
        # rolled back by switchStateV(0); the converged states are clean)
        assert np.array_equal(new[k], old[k], equal_nan=old[k].dtype.kind == "f"), k
    assert not np.isnan(old["s2.end.cp_A_single"]).any() and not np.isnan(old["s2.end.F"]).any()
    assert old["setup.neighbors"].shape[1] == 14 and old["setup.conn"].shape[1] == 41
    assert int(old["s1.end.cp_Jact"].sum()) > 0 and int(old["s2.end.cp_Jact"].sum()) > 0


def test_golden_internal_consistency(golden):
    g = golden
    assert g["setup.xyz"].shape == (216, 3)
    assert list(g["newton_counts"]) == [36, 49]
    # residual = dispBC_index * (Pex - Pin)  (stiffness.c:527)
    Pin = g["s1.pred.Pin"].reshape(-1, 3)
    res = g["s1.bc.dispBC_index"] * (g["s1.bc.Pex"] - Pin.reshape(-1))
    assert np.array_equal(res, g["s1.rr.residual"])
    # IK/JK layout of SURVEY Appendix B
    kp, IK = g["setup.K_pointer"], g["s1.fd.IK"]
    assert np.array_equal(IK[0::3][:-1], kp[:-1, 1] + 1)
    assert np.array_equal(IK[1::3], kp[:-1, 1] + 3 * kp[:-1, 0] + 1)
    assert np.array_equal(IK[2::3], kp[:-1, 1] + 6 * kp[:-1, 0])
    assert IK[-1] == kp[-1, 1] + 1


def test_reference_rounding_noise_floor(tmp_path):
    """How reproducible is the REFERENCE itself?  Run its first three default load steps twice, changing
    only the summation order inside the open ddot (naive loop vs pairwise; MKL's real order is unknown):
    displacements agree to ~1e-10 but the step-3 bond forces move by ~1e-8.  This is the noise floor that
    bounds any trajectory-level parity claim (SURVEY section 7 hard part 1, section 6 thread sensitivity)."""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    import subprocess, sys, os
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle.ref import RefLPM\nimport numpy as np\n"
        "r = RefLPM.instance(); r.threads(1); r.setup_sc()\n"
        "out = {}\n"
        "for s in (1, 2, 3):\n"
        "    r.load_step(s, [(1, 'z', 0.0)], [(2, 0.0, 0.0, -2000.0)])\n"
        "    out['F%%d' %% s] = r.get('F'); out['x%%d' %% s] = r.get('xyz') - r.get('xyz_initial')\n"
        "np.savez(sys.argv[1], **out)\n" % str(root))
    procs = []
    for mode in ("naive", "pairwise"):
        env = dict(os.environ, LPMB_SHIM_DOT=mode, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", code, str(tmp_path / f"{mode}.npz")], env=env,
                                      stdout=subprocess.DEVNULL))
    assert all(p.wait() == 0 for p in procs)
    a, b = np.load(tmp_path / "naive.npz"), np.load(tmp_path / "pairwise.npz")
    rel = lambda k: float(np.linalg.norm(a[k] - b[k]) / np.linalg.norm(b[k]))
    assert rel("x1") < 1e-12 and rel("F1") < 1e-10 and rel("F2") < 1e-9
    assert rel("x3") < 1e-9
    assert 1e-9 < rel("F3") < 1e-7          # the reference's own bond forces are not reproducible to 1e-9 here


@pytest.mark.parametrize("box", [(-0.2, 10.2, -0.2, 10.2, -0.2, 10.2), (-0.2, 3.2, -0.2, 5.7, -0.2, 2.2)])
def test_injected_lattice_topology_equals_the_reference_search(ref, box):
    """oracle/ref.py::inject_sc_topology (O(N) lattice stencils in numpy; what lets bench.py's reference arm reach S1 = 100^3)
    leaves exactly what searchNormalNeighbor() + searchAFEMNeighbor() (neighbor.c:9-141, O(N^2)) leave in the reference's
    globals: neighbour lists in ascending-j order with interleaved shells, shell signs, counts, initial distances and unit
    vectors bit for bit, conn / nb_conn / K_pointer and the CSR size -- on the default 21^3 block and a ragged 7 x 12 x 5 one"""
    r = ref
    names = ["neighbors", "nsign", "nb", "nb_initial", "conn", "nb_conn", "K_pointer", "distance_initial", "csx_initial", "csy_initial",
             "csz_initial"]
    r.setup_sc(box=box)
    a = {n: r.get(n) for n in names}
    a["neighbors1"], a["neighbors2"] = r.i2("neighbors1", r.N, 6), r.i2("neighbors2", r.N, 12)
    r.setup_sc(box=box, neighbor_search="lattice")
    b = {n: r.get(n) for n in names}
    b["neighbors1"], b["neighbors2"] = r.i2("neighbors1", r.N, 6), r.i2("neighbors2", r.N, 12)
    for n in a:
        assert np.array_equal(a[n], b[n]), n
    assert int(b["K_pointer"][r.N, 1]) == int(a["K_pointer"][r.N, 1]) > 0
