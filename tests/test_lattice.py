"""CPU: the O(N) stencil builders reproduce the reference's O(N^2) neighbour search bit for bit."""
import numpy as np


def test_sc_block_matches_reference_default_case(lpm, ref_c1):
    r = ref_c1["ref"]
    lat = lpm.lattice.sc_block(21)
    assert np.array_equal(lat["neighbors"], r.get("neighbors"))
    assert np.array_equal(lat["nsign"], r.get("nsign"))
    assert np.array_equal(lat["nb"], r.get("nb_initial"))
    assert np.array_equal(lat["conn"], r.get("conn"))
    assert np.array_equal(lat["nb_conn"], r.get("nb_conn"))
    kp = lpm.lattice.k_pointer(lat["conn"], 3)
    assert np.array_equal(kp, r.get("K_pointer").astype(np.int64))


def test_sc_block_matches_golden(lpm, golden):
    lat = lpm.lattice.sc_block(6)
    assert np.array_equal(lat["neighbors"], golden["setup.neighbors"])
    assert np.array_equal(lat["nsign"], golden["setup.nsign"])
    assert np.array_equal(lat["conn"], golden["setup.conn"])


def test_sizes_s1(lpm):
    """S1 = SC 100^3: sum(nb_conn)=59 157 952, nnz_upper=267 710 784, from the stencil
    counts without materialising the lists"""
    first, second, conn = lpm.lattice.sc_offsets()
    assert len(first) == 6 and len(second) == 12 and len(conn) == 61
    n = 100
    nblk = sum((n - abs(a)) * (n - abs(b)) * (n - abs(c)) for a, b, c in conn)
    assert nblk == 59157952
    upper = (nblk - n ** 3) // 2 + n ** 3          # blocks with column >= row
    assert 9 * upper - 3 * n ** 3 == 267710784
