"""GPU parity: stiffness container, SpMV and CG (lpmb_solver.cu) against the oracle.

Reference path: solverCG(), src/solver.c:188-270, on the symmetric-upper 1-based CSR that
src/stiffness.c:441-515 fills.  Tolerances (fp64): container round trip bit-exact; SpMV 1e-13
relative; CG same iteration count and disp within 1e-10 relative (north_star asks 1e-9).
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def full_from_upper(K, IK, JK, n):
    U = sp.csr_matrix((K, JK - 1, IK - 1), shape=(n, n))
    return (U + sp.triu(U, 1).T).tocsr()


@pytest.fixture(scope="module")
def ctx6(lpm, golden):
    c = lpm.Context(216, 3, 2, 18, 61)
    c.set_connectivity(golden["setup.conn"])
    yield c
    c.close()


def test_csr_layout_bit_exact(ctx6, golden):
    """K_pointer / IK / JK reproduce neighbor.c:114-130 + stiffness.c:486-515 exactly"""
    nnz, nblk = ctx6.csr_sizes()
    assert nnz == int(golden["setup.K_pointer"][-1, 1])
    assert nblk == int(golden["setup.nb_conn"].sum())
    assert np.array_equal(ctx6.k_pointer(), golden["setup.K_pointer"])
    ctx6.matrix_from_upper_csr(golden["s1.fd.K_global"])
    K, IK, JK = ctx6.matrix_to_upper_csr()
    assert np.array_equal(IK, golden["s1.fd.IK"])
    assert np.array_equal(JK, golden["s1.fd.JK"])
    assert np.array_equal(K, golden["s1.fd.K_global"])  # import -> SELL -> export is lossless


def test_spmv_matches_reference_matrix(ctx6, golden):
    Kbc = golden["s1.n0.K_bc"]
    ctx6.matrix_from_upper_csr(Kbc)
    A = full_from_upper(Kbc, golden["s1.fd.IK"], golden["s1.fd.JK"], 648)
    rng = np.random.default_rng(20240607)
    for _ in range(3):
        x = rng.standard_normal(648)
        y = ctx6.spmv(x)
        yr = A @ x
        assert np.abs(y - yr).max() <= 1e-13 * np.abs(yr).max()


@pytest.mark.parametrize("tag", ["s1.n0", "s1.n1", "s1.n2", "s2.n0"])
def test_cg_matches_reference_solve(ctx6, golden, tag):
    """same K (BC-modified by the reference's boundary.c), same rhs -> same iteration count, same disp"""
    ctx6.matrix_from_upper_csr(golden[f"{tag}.K_bc"])
    x, iters, ok = ctx6.solve_cg(golden[f"{tag}.rhs"])
    assert ok
    assert iters == int(golden[f"{tag}.cg_iters"][0])
    ref = golden[f"{tag}.disp"]
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)


def test_masked_cg_equals_bc_modified_matrix(ctx6, golden):
    """DoF masking inside the solve == the reference's row/column zeroing + norm_diag on the diagonal
    (boundary.c:159-281): constrained DoFs never enter the Krylov space (rhs 0, x0 = 0)."""
    ctx6.matrix_from_upper_csr(golden["s1.fd.K_global"])          # un-modified tangent
    ctx6.set_dof_mask(golden["s1.bc.dispBC_index"], golden["s1.bc.fix_index"])
    x, iters, ok = ctx6.solve_cg(golden["s1.rr.residual"], use_mask=True)
    assert ok and iters == int(golden["s1.n0.cg_iters"][0])
    ref = golden["s1.n0.disp"]
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
    assert np.all(x[golden["s1.bc.dispBC_index"] == 0] == 0.0)


def test_graph_replayed_cg_batches_are_bit_identical(ctx6, golden):
    """small lattices: the 16 iterations between two convergence polls are replayed as one CUDA graph (param cg_graph, default
    on) -- same kernels, same order: iteration counts and every bit of the solution equal the plain-launch path, also when the
    graph is re-captured for another mask / maxit and when a solve ends inside a batch"""
    ctx6.matrix_from_upper_csr(golden["s1.n0.K_bc"])
    res = {}
    for graph in (1.0, 0.0, 1.0):
        ctx6.set_params(cg_graph=graph)
        res.setdefault(graph, []).append(ctx6.solve_cg(golden["s1.n0.rhs"]))
        res[graph].append(ctx6.solve_cg(golden["s1.n0.rhs"], maxit=19))   # 23 needed: one graph batch + 3 plain iterations
    ctx6.matrix_from_upper_csr(golden["s1.fd.K_global"])
    ctx6.set_dof_mask(golden["s1.bc.dispBC_index"], golden["s1.bc.fix_index"])
    for graph in (1.0, 0.0):
        ctx6.set_params(cg_graph=graph)
        res[graph].append(ctx6.solve_cg(golden["s1.rr.residual"], use_mask=True))
    assert res[1.0][0][1] == int(golden["s1.n0.cg_iters"][0]) and res[1.0][0][2]
    assert res[1.0][1][1] == 19 and not res[1.0][1][2]
    plain = res[0.0]
    for k, (x, it, ok) in enumerate(res[1.0][:2] + res[1.0][4:5]):
        assert it == plain[k][1] and ok == plain[k][2] and np.array_equal(x, plain[k][0]), k
    for k in (2, 3):   # the second pass with the graph (re-used executable)
        assert np.array_equal(res[1.0][k][0], plain[k - 2][0]) and res[1.0][k][1] == plain[k - 2][1]


def test_cg_zero_rhs_and_maxit(ctx6, golden):
    ctx6.matrix_from_upper_csr(golden["s1.n0.K_bc"])
    x, iters, ok = ctx6.solve_cg(np.zeros(648))
    assert ok and iters == 0 and not x.any()
    x, iters, ok = ctx6.solve_cg(golden["s1.n0.rhs"], maxit=5)
    assert not ok and iters == 5


def test_default_case_solves_like_reference(lpm, ref_c1):
    """C1 (21^3, n=27 783, nnz_upper=2 203 713): 80 then 106 CG iterations"""
    r = ref_c1["ref"]
    L = r.lib
    c = lpm.Context(r.N, 3, 2, 18, 61)
    c.set_connectivity(r.get("conn"))
    for expect in (80, 106):
        L.switchStateV(0)
        L.setDispBC_stiffnessUpdate3D()
        Kbc, rhs = r.get("K_global"), r.get("residual")
        c.matrix_from_upper_csr(Kbc)
        x, iters, ok = c.solve_cg(rhs)
        L.solverCG()
        assert ok and iters == expect == L.lpmb_shim_last_itercount()
        ref = r.get("disp")
        assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
        L.computeBondForceGeneral(0, 1)
        L.updateRR()
    c.close()


def test_large_lattice_properties(lpm):
    """size-independent properties at 64^3 (262 144 particles): symmetry x.(Ky) = y.(Kx), linearity,
    and CG residual ||b - Kx|| <= 1.01e-4 ||b|| on the SPD test pattern"""
    lat = lpm.lattice.sc_block(64)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_connectivity(lat["conn"])
    nnz, nblk = c.csr_sizes()
    assert nblk == int(lat["nb_conn"].sum())
    assert nnz == int(lpm.lattice.k_pointer(lat["conn"], 3)[-1, 1])
    c.fill_test_pattern()
    rng = np.random.default_rng(20240607)
    x, y = rng.standard_normal(3 * N), rng.standard_normal(3 * N)
    Kx, Ky = c.spmv(x), c.spmv(y)
    assert abs(x @ Ky - y @ Kx) <= 1e-12 * abs(x @ Ky)
    Kxy = c.spmv(2.0 * x - 3.0 * y)
    assert np.abs(Kxy - (2.0 * Kx - 3.0 * Ky)).max() <= 1e-11 * np.abs(Kxy).max()
    b = rng.standard_normal(3 * N)
    sol, iters, ok = c.solve_cg(b)
    assert ok and 0 < iters < 200
    r = b - c.spmv(sol)
    assert np.linalg.norm(r) <= 1.01e-4 * np.linalg.norm(b)
    c.close()


@pytest.mark.parametrize("dims", [(8, 8, 8), (20, 13, 9), (17, 24, 33)])
def test_brick_spmv_matches_full_format(lpm, dims):
    """brick-blocked symmetric kernel (lpmb_brick.cu) == full-format SELL kernel, including partial bricks,
    multi-brick staging in every direction and the masked CG"""
    lat = lpm.lattice.sc_block(*dims, h=0.5, origin=(0.3, -1.7, 2.0))
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25)
    c.set_field("xyz_initial", lat["xyz"])
    c.set_connectivity(lat["conn"])
    c.fill_test_pattern()
    rng = np.random.default_rng(20240607)
    x = rng.standard_normal(3 * N)
    b = rng.standard_normal(3 * N)
    y0 = c.spmv(x)
    bc = np.ones(3 * N, dtype=np.int32)
    bc[rng.integers(0, 3 * N, size=N // 10)] = 0
    c.set_dof_mask(bc, np.ones(3 * N, dtype=np.int32))
    d0, it0, ok0 = c.solve_cg(b * bc, use_mask=True)
    c.enable_bricks(True)
    y1 = c.spmv(x)
    assert np.abs(y0 - y1).max() <= 1e-13 * np.abs(y0).max()
    d1, it1, ok1 = c.solve_cg(b * bc, use_mask=True)
    assert ok0 and ok1 and abs(it0 - it1) <= 1
    assert np.linalg.norm(d0 - d1) <= 1e-9 * np.linalg.norm(d0)
    assert np.all(d1[bc == 0] == 0.0)
    # values follow the matrix: re-import changes the product
    c.enable_bricks(False)
    y2 = c.spmv(x)
    assert np.array_equal(y0, y2)
    c.close()


@pytest.mark.parametrize("dims,own", [((16, 12, 31), (2, 29)), ((9, 17, 14), (0, 12)), ((12, 8, 14), (2, 14)), ((8, 8, 21), (0, 21))])
def test_brick_tiles_stream_only_needed_rows(lpm, dims, own):
    """Only the z-layers of a class tile whose rows are needed are streamed (lpmb_brick.cu, `zr` table): empty rows of a
    partial brick layer, the upper CG-halo rows of a slab (they own no pair touching an owned row) and, per class, the
    lower-halo rows that do not reach into the slab.  A slab is emulated on one GPU with the test hook
    brick_own_z0 / brick_own_z1: x is arbitrary on the 'halo' layers (as after a halo push), the products of the OWNED
    rows must equal the full-format kernel's, and the masked CG must behave the same.  (2, 29) of 31 layers is the
    216^3 / 8-GPU slab shape; (0, 12) / (2, 14) of 14 are the two slabs of tests/dist_check.py."""
    lat = lpm.lattice.sc_block(*dims, h=0.5, origin=(0.3, -1.7, 2.0))
    N = lat["xyz"].shape[0]
    nx, ny, nz = dims
    z = np.arange(N) // (nx * ny)
    owned = (z >= own[0]) & (z < own[1])
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25, brick_own_z0=own[0], brick_own_z1=own[1])
    c.set_field("xyz_initial", lat["xyz"])
    c.set_connectivity(lat["conn"])
    c.fill_test_pattern()
    rng = np.random.default_rng(7)
    x = rng.standard_normal(3 * N)
    b = rng.standard_normal(3 * N)
    y0 = c.spmv(x).reshape(N, 3)
    bc = np.repeat(owned.astype(np.int32), 3)                # "ghost" DoFs are masked like lpmb_refresh_mask does in slab runs
    bc[rng.integers(0, 3 * N, size=N // 10)] = 0
    c.set_dof_mask(bc, np.ones(3 * N, dtype=np.int32))
    d0, it0, ok0 = c.solve_cg(b * bc, use_mask=True)
    full_bytes = None
    for trim in (0.0, 1.0):
        c.set_params(brick_trim=trim)
        c.enable_bricks(True)
        y1 = c.spmv(x).reshape(N, 3)
        assert np.abs(y0[owned] - y1[owned]).max() <= 1e-13 * np.abs(y0).max(), trim
        d1, it1, ok1 = c.solve_cg(b * bc, use_mask=True)
        assert ok0 and ok1 and abs(it0 - it1) <= 1
        assert np.linalg.norm(d0 - d1) <= 1e-9 * np.linalg.norm(d0)
        nbytes = c.spmv_bytes_bricks()
        if trim == 0.0:
            full_bytes = nbytes
        else:
            assert nbytes < full_bytes                         # every case here has rows to skip
        c.enable_bricks(False)
    c.close()


def test_brick_cg_on_reference_tangent(lpm, golden):
    """FD tangent of the golden 6^3 case through the brick kernel: reference iteration count and disp"""
    from helpers import make_ctx
    c = make_ctx(lpm, golden)
    c.fd_stiffness(False)
    c.set_dof_mask(golden["s1.bc.dispBC_index"], golden["s1.bc.fix_index"])
    c.enable_bricks(True)
    x, iters, ok = c.solve_cg(golden["s1.rr.residual"], use_mask=True)
    assert ok and iters == int(golden["s1.n0.cg_iters"][0])
    ref = golden["s1.n0.disp"]
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
    # a new assembly invalidates the brick values; the next solve refreshes them
    c.matrix_from_upper_csr(golden["s1.fd.K_global"])
    x2, iters2, ok2 = c.solve_cg(golden["s1.rr.residual"], use_mask=True)
    assert ok2 and iters2 == iters and np.array_equal(x, x2)
    c.close()


def test_brick_rejects_unsupported_lattice(lpm, golden):
    lat = lpm.lattice.sc_block(6)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25)
    xyz = lat["xyz"].copy()
    xyz[5, 0] += 0.1          # off-lattice particle
    c.set_field("xyz_initial", xyz)
    c.set_connectivity(lat["conn"])
    with pytest.raises(RuntimeError):
        c.enable_bricks(True)
    c.fill_test_pattern()
    x = np.ones(3 * N)
    assert np.isfinite(c.spmv(x)).all()   # full format still in use
    c.close()


def test_brick_spmv_on_carved_lattice(lpm):
    """irregular specimen (notch + hole carved out of an SC block, like the reference's CT geometry): topology from
    the O(N) device builder, bricks partially filled, rows with short conn lists"""
    lat = lpm.lattice.sc_block(19, 14, 11, h=0.5)
    xyz = lat["xyz"]
    keep = ~((xyz[:, 0] < 4.0) & (np.abs(xyz[:, 1] - 3.25) < 0.4))                      # notch through the thickness
    keep &= ((xyz[:, 0] - 6.5) ** 2 + (xyz[:, 1] - 4.0) ** 2) > 1.2 ** 2                # pin hole
    xyz = np.ascontiguousarray(xyz[keep])
    N = xyz.shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25)
    c.set_field("xyz", xyz)
    c.set_field("xyz_initial", xyz)
    c.build_topology(0.5, np.sqrt(2.0) * 0.5)
    c.fill_test_pattern()
    rng = np.random.default_rng(20240607)
    x = rng.standard_normal(3 * N)
    y0 = c.spmv(x)
    c.enable_bricks(True)
    y1 = c.spmv(x)
    assert np.abs(y0 - y1).max() <= 1e-13 * np.abs(y0).max()
    b = rng.standard_normal(3 * N)
    d1, it1, ok1 = c.solve_cg(b)
    c.enable_bricks(False)
    d0, it0, ok0 = c.solve_cg(b)
    assert ok0 and ok1 and abs(it0 - it1) <= 1 and np.linalg.norm(d0 - d1) <= 1e-9 * np.linalg.norm(d0)
    c.close()


@pytest.mark.parametrize("bricks", [False, True])
def test_preconditioned_fast_mode(lpm, bricks):
    """Opt-in fast mode (param cg_precond = 1; lpmb_mg.cu + pcg_run): CG preconditioned with the matrix-free multigrid V-cycle
    on the bench workload at 24^3 (C5 material, 1 % stretch, top / bottom layers constrained in z).  Not the parity path --
    the reference's solverCG is unpreconditioned (solver.c:219-220) -- so it is validated the way SURVEY section 7 asks:
    (1) same stop rule on the TRUE residual: ||mask (b - K x)|| <= 1e-4 ||mask b||; (2) at least 3x fewer iterations than the
    plain CG (numpy prototype: 11 vs 64); (3) both modes run to a relative residual of 1e-11 agree to 1e-8 (bottom layer clamped so that K is non-singular); (4) switching the mode off restores the
    parity iteration count exactly."""
    import bench
    n = 24
    c, info = bench.build_workload(lpm, n, 0, bricks=bricks)
    N = n ** 3
    # the bench workload holds the end layers in z only: K is singular (rigid x / y translations, rotation about z), two
    # Krylov methods then differ by a null-space vector.  Clamp the bottom layer completely for the comparison of solutions.
    c.apply_disp_bc(2, "x", 0.0)
    c.apply_disp_bc(2, "y", 0.0)
    bc, fix = c.get_field("dispBC_index"), c.get_field("fix_index")
    c.set_dof_mask(bc, fix)
    mask = ((bc != 0) & (fix != 0)).astype(float)
    b = c.get_field("residual") * mask

    def solve(rel):
        c.copy_field("residual", "residual_save")
        it, ok = c.solve_cg_device(rel=rel, abs_tol=1e-12 if rel > 1e-20 else 0.0, update_xyz=False)
        return it, ok, c.get_field("disp")

    it0, ok0, d0 = solve(1e-8)
    c.set_params(cg_precond=1.0)
    it1, ok1, d1 = solve(1e-8)
    it1t, ok1t, d1t = solve(1e-22)
    c.set_params(cg_precond=0.0)
    it0t, ok0t, d0t = solve(1e-22)
    it0b, _, d0b = solve(1e-8)
    assert ok0 and ok1 and ok0t and ok1t
    assert it1 * 3 <= it0, (it1, it0)
    # both enabled kernels see the same operator: use the full-format host SpMV for the true residual
    r1 = mask * (b - c.spmv(d1))
    assert np.linalg.norm(r1) <= 1.0001e-4 * np.linalg.norm(b), (np.linalg.norm(r1), np.linalg.norm(b))
    assert np.abs(d1[mask == 0]).max() == 0.0
    assert np.linalg.norm(d1t - d0t) <= 1e-8 * np.linalg.norm(d0t), np.linalg.norm(d1t - d0t) / np.linalg.norm(d0t)
    assert it0b == it0 and np.array_equal(d0b, d0)
    print(f"fast mode at {n}^3 (bricks={bricks}): {it1} PCG iterations vs {it0} CG iterations; to 1e-11: {it1t} vs {it0t}")
    c.close()


def test_fast_mode_tiled_stencil_kernel_equals_plain_kernel(lpm):
    """the shared-memory tiled multigrid stencil kernel (used from 32^3 sites per level) forced on at 24^3 (param
    mg_tiled_min = 0, so that partial tiles and boundary tiles on every level are exercised): same PCG iteration count and the
    same displacement as the plain kernel -- both evaluate the identical sequence of fused multiply-adds per site"""
    import bench
    c, info = bench.build_workload(lpm, 24, 0, bricks=False)
    c.set_dof_mask(c.get_field("dispBC_index"), c.get_field("fix_index"))
    c.set_params(cg_precond=1.0)
    res = []
    for tiled_min in (1e18, 0.0):
        c.set_params(mg_tiled_min=tiled_min)
        c.copy_field("residual", "residual_save")
        it, ok = c.solve_cg_device(update_xyz=False)
        res.append((it, ok, c.get_field("disp")))
    assert res[0][1] and res[1][1] and res[0][0] == res[1][0], (res[0][0], res[1][0])
    assert np.abs(res[0][2] - res[1][2]).max() <= 1e-13 * np.abs(res[0][2]).max()
    c.close()


def test_fast_mode_two_contexts_interleaved(lpm):
    """the multigrid stencil sits in __constant__ memory, i.e. once per device: two lattices with DIFFERENT tangents solved
    alternately in one process must each see their own stencil (the context that uploaded last is tracked and the other one
    re-uploads): the second solve of context A equals its first bit for bit although context B solved in between"""
    import bench
    a, _ = bench.build_workload(lpm, 24, 0, bricks=False)
    saved = dict(bench.PHYS)
    try:
        bench.PHYS.update(E0=210e3, mu0=0.2)            # another material -> another 61-point stencil
        b, _ = bench.build_workload(lpm, 20, 0, bricks=False)
    finally:
        bench.PHYS.clear()
        bench.PHYS.update(saved)
    out = []
    for c in (a, b):
        c.set_dof_mask(c.get_field("dispBC_index"), c.get_field("fix_index"))
        c.set_params(cg_precond=1.0)

    def solve(c):
        c.copy_field("residual", "residual_save")
        it, ok = c.solve_cg_device(update_xyz=False)
        assert ok
        return it, c.get_field("disp")

    ia1, da1 = solve(a)
    ib1, db1 = solve(b)
    ia2, da2 = solve(a)
    ib2, db2 = solve(b)
    assert ia1 == ia2 and np.array_equal(da1, da2), (ia1, ia2)
    assert ib1 == ib2 and np.array_equal(db1, db2), (ib1, ib2)
    assert ia1 <= 15 and ib1 <= 15, (ia1, ib1)
    a.close()
    b.close()


def test_fast_mode_fails_loudly_when_not_eligible(lpm, golden):
    """cg_precond = 1 on a lattice the multigrid hierarchy does not cover (here: a block only 4 sites thick, so no interior
    61-point stencil row exists) must return an error that names the parameter -- no silent fall back to another solver"""
    lat = lpm.lattice.sc_block(6, 6, 4, h=0.5)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25, cg_precond=1.0)
    c.set_field("xyz_initial", lat["xyz"])
    c.set_connectivity(lat["conn"])
    c.fill_test_pattern()
    with pytest.raises(Exception) as e:
        c.solve_cg(np.ones(3 * N))
    assert "cg_precond" in str(e.value)
    c.close()
