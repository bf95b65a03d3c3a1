"""GPU parity of the slice-cooperative FD assembly (`fd_shell_table_kernel` + `fd_rows_kernel`, lpmb_stiffness.cu; param
`fd_variant` = 1, the default) beyond the sizes the golden vectors cover.

The golden-vector tests (`test_constitutive_gpu.py::test_fd_stiffness_bit_exact`, `test_variants_gpu.py`, the drop-in
runs) pin the default variant on the reference's own `K_global`.  Here the CTA-per-particle kernel of round 1
(`fd_variant = 0`, itself bit-exact on those fixtures) is the second witness on lattices the oracle would need minutes
for: deformed, plastically pre-stretched and damaged blocks whose SELL slices straddle lattice rows and free surfaces,
and the FCC block of BASELINE config 4.  Both variants evaluate the reference's expressions (stiffness.c:384-516,
constitutive.c:228-283) in its order, so every stored value, every side effect (F, Pin of the last perturbation; dL, cs*,
dL_total, TdL_total of the last toucher) must agree BIT FOR BIT.
"""
import numpy as np
import pytest

from helpers import assert_same

pytestmark = pytest.mark.gpu

RADIUS = 0.25


def _block(lpm, lattice, nx, ny, nz, seed):
    """undamaged block of lattice 2 (simple cubic, 18 neighbours / 61 conn) or 3 (FCC, 12 + 6 neighbours) with isotropic Kn / Tv"""
    h = 2 * RADIUS
    rng = np.random.default_rng(seed)
    if lattice == 2:
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        xyz = h * np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
        cut1, cut2, nconn = h, np.sqrt(2.0) * h, 61
    else:
        a = 2.0 * np.sqrt(2.0) * RADIUS   # FCC cell edge; nearest neighbours at 2 radius, second shell at a
        pts = []
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    for b in ((0, 0, 0), (0.5, 0.5, 0), (0.5, 0, 0.5), (0, 0.5, 0.5)):
                        pts.append(((i + b[0]) * a, (j + b[1]) * a, (k + b[2]) * a))
        xyz = np.array(pts)
        xyz = xyz[np.lexsort((xyz[:, 0], xyz[:, 1], xyz[:, 2]))]
        cut1, cut2, nconn = 2 * RADIUS, a, 61
    N = len(xyz)
    c = lpm.Context(N, 3, lattice, 18, nconn)
    c.set_params(radius=RADIUS, particle_volume=h ** 3)
    c.set_field("xyz", xyz)
    c.set_field("xyz_initial", xyz)
    c.build_topology(cut1, cut2)   # the reference's cut-offs (tolerance inside, neighbor.c:17-21)
    c.set_field("type", np.zeros(N, dtype=np.int32))
    E0, mu0 = 69e3, 0.3
    C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0)
    C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0)
    C44 = E0 / 2.0 / (1.0 + mu0) * (1.0 if lattice == 2 else 1.3)
    c.calc_kntv(np.tile([C11, C12, C44], (3, 1)))
    c.compute_dl()
    # deformed configuration, committed plastic stretch, broken bonds (both directions of a bond, as updateCrack leaves them)
    c.set_field("xyz", xyz + rng.normal(scale=2e-3 * h, size=xyz.shape))
    nbr = c.get_field("neighbors")
    mir = c.get_field("mirror")
    dlp = rng.normal(scale=1e-4 * h, size=nbr.shape) * (rng.random(nbr.shape) < 0.3)
    dlp[nbr < 0] = 0.0
    c.set_field("dLp0", dlp)
    brk = np.ones(nbr.shape)
    ii, jj = np.nonzero((nbr >= 0) & (rng.random(nbr.shape) < 0.02))
    brk[ii, jj] = 0.0
    brk[nbr[ii, jj], mir[ii, jj]] = 0.0
    # a one-sided break as well (the law multiplies by the row's own flag only)
    brk[N // 2, 0] = 0.0
    c.set_field("damage_broken", brk)
    return c


def _assemble(c, variant):
    c.set_params(fd_variant=float(variant))
    c.fd_stiffness(True)
    K, IK, JK = c.matrix_to_upper_csr()
    side = {n: c.get_field(n) for n in ("F", "Pin", "dL", "csx", "csy", "csz", "dL_total", "TdL_total")}
    return K, IK, JK, side


@pytest.mark.parametrize("lattice,shape", [(2, (13, 11, 9)), (2, (37, 5, 6)), (3, (5, 4, 3))])
def test_slice_cooperative_assembly_equals_cta_per_particle_kernel(lpm, lattice, shape):
    c = _block(lpm, lattice, *shape, seed=7 + lattice)
    K0, IK0, JK0, s0 = _assemble(c, 0)
    assert np.isfinite(K0).all() and np.abs(K0).max() > 0
    # wipe the values in between so that a kernel that writes nothing cannot pass
    c.fill_test_pattern()
    K1, IK1, JK1, s1 = _assemble(c, 1)
    assert np.array_equal(IK0, IK1) and np.array_equal(JK0, JK1)
    assert_same(K1, K0, "K_global (slice-cooperative vs CTA-per-particle)")
    for n in s0:
        assert_same(s1[n], s0[n], f"side effect {n}")
    c.close()


def test_assembly_is_repeatable_and_default_is_the_new_variant(lpm):
    c = _block(lpm, 2, 9, 9, 9, seed=3)
    c.fd_stiffness(True)                     # default variant
    Kd = c.matrix_to_upper_csr()[0]
    K1 = _assemble(c, 1)[0]
    K0 = _assemble(c, 0)[0]
    assert_same(Kd, K1, "default == variant 1")
    assert_same(K1, K0, "variant 1 == variant 0")
    c.close()


def test_chunked_table_window(lpm):
    """the table of perturbed shell sums is bounded scratch (param fd_tab_mb): with a budget far below the lattice the rows
    are assembled chunk by chunk against a sliding window of the table -- same bits"""
    c = _block(lpm, 2, 37, 5, 6, seed=11)      # band = 37 * 5 + 37 + 1 = 223 particles
    K0, _, _, s0 = _assemble(c, 0)
    c.fill_test_pattern()
    c.set_params(fd_tab_mb=0.7)                # 736 particles in the table -> 288 rows per chunk, 4 chunks
    K1, _, _, s1 = _assemble(c, 1)
    assert_same(K1, K0, "K_global, chunked")
    for n in s0:
        assert_same(s1[n], s0[n], f"side effect {n}, chunked")
    c.set_params(fd_tab_mb=2048.0)
    c.fill_test_pattern()
    K2 = _assemble(c, 1)[0]
    assert_same(K2, K0, "K_global, one chunk after the table grew")
    c.close()
