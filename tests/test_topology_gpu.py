"""GPU parity: O(N) device neighbour search / AFEM connectivity vs the reference's O(N^2) search
(neighbor.c:9-141).  Integer outputs bit-exact; distance_initial / cs*_initial bit-exact."""
import numpy as np
import pytest

from helpers import assert_same, params_from_golden

pytestmark = pytest.mark.gpu


def test_build_topology_matches_golden(lpm, golden):
    g = golden
    p = params_from_golden(g)
    c = lpm.Context(216, 3, 2, 18, 61)
    c.set_field("xyz", g["setup.xyz"])
    c.build_topology(p["neighbor1_cutoff"], p["neighbor2_cutoff"])
    assert_same(c.get_field("neighbors"), g["setup.neighbors"], "neighbors")
    assert_same(c.get_field("nsign"), g["setup.nsign"], "nsign")
    assert_same(c.get_field("nb_initial"), g["setup.nb_initial"], "nb_initial")
    for n in ("distance_initial", "csx_initial", "csy_initial", "csz_initial"):
        assert_same(c.get_field(n), g[f"setup.{n}"], n)
    assert np.array_equal(c.k_pointer(), g["setup.K_pointer"])
    nnz, nblk = c.csr_sizes()
    assert nblk == int(g["setup.nb_conn"].sum())
    # conn itself: export the pattern through JK of an (all-zero) matrix
    c.fill_test_pattern()
    _, IK, JK = c.matrix_to_upper_csr()
    assert np.array_equal(IK, g["s1.fd.IK"]) and np.array_equal(JK, g["s1.fd.JK"])
    c.close()


def test_build_topology_default_case(lpm, ref_c1):
    """C1: 21^3 with the reference's own coordinates (origin quirk of initialization.c:269-275 included)"""
    r = ref_c1["ref"]
    c = lpm.Context(r.N, 3, 2, 18, 61)
    c.set_field("xyz", r.get("xyz_initial"))
    c.build_topology(r.gd("neighbor1_cutoff"), r.gd("neighbor2_cutoff"))
    assert_same(c.get_field("neighbors"), r.get("neighbors"), "neighbors")
    assert_same(c.get_field("nsign"), r.get("nsign"), "nsign")
    assert_same(c.get_field("distance_initial"), r.get("distance_initial"), "distance_initial")
    assert np.array_equal(c.k_pointer(), r.get("K_pointer"))
    c.close()


def test_build_topology_large_block_vs_stencil(lpm):
    """48^3 = 110 592 particles: device cell-list search == closed-form lattice stencil"""
    lat = lpm.lattice.sc_block(48)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_field("xyz", lat["xyz"])
    c.build_topology(0.5, 0.5 * np.sqrt(2.0))
    assert_same(c.get_field("neighbors"), lat["neighbors"], "neighbors")
    assert_same(c.get_field("nsign"), lat["nsign"], "nsign")
    assert np.array_equal(c.k_pointer().astype(np.int64), lpm.lattice.k_pointer(lat["conn"], 3))
    c.close()


def test_ragged_and_empty_inputs(lpm):
    """edge cases: an isolated particle far from a small cluster, and neighbour overflow detection"""
    xyz = np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.5, 0], [10, 10, 10]], dtype=np.float64)
    c = lpm.Context(4, 3, 2, 18, 61)
    c.set_field("xyz", xyz)
    c.build_topology(0.5, 0.5 * np.sqrt(2.0))
    nbr = c.get_field("neighbors")
    assert list(nbr[3]) == [-1] * 18 and list(nbr[0][:3]) == [1, 2, -1]
    assert list(c.get_field("nsign")[1][:2]) == [0, 1]
    c.close()
    # 40 coincident-ish particles exceed nneighbors=18: the reference would overrun its arrays; we refuse
    rng = np.random.default_rng(20240607)
    c = lpm.Context(40, 3, 2, 18, 61)
    c.set_field("xyz", 0.01 * rng.standard_normal((40, 3)))
    with pytest.raises(lpm.LPMBError):
        c.build_topology(0.5, 0.5 * np.sqrt(2.0))
    c.close()
