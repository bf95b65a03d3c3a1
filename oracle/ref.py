"""oracle/ref.py -- drive the UNMODIFIED reference (oracle/_ref/liblpmc_ref.so) from Python.

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product path.

liblpmc_ref.so is the reference's own src/*.c compiled as they lie in /root/reference
(oracle/Makefile) against the open MKL stand-in oracle/shim/.  The default driver's main()
is renamed at compile time, so the ~140 process globals it defines
(src/lpmc_project.c:19-45, declared in include/lpm.h:55-81) are plain exported data symbols
that ctypes can read and write, and every library function (stiffness.h, solver.h,
constitutive.h, neighbor.h, boundary.h, initialization.h, lpm_basic.h) can be called.

`RefLPM.setup_sc()` re-plays the set-up section of the default driver
(src/lpmc_project.c:75-345) with the box / material as parameters; `load_step()` re-plays
one pass of its load-step loop (src/lpmc_project.c:382-546), optionally with hooks so a
test can substitute single entry points.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
# LPMB_REF_SO selects another build of the same driver TU, e.g. _ref/liblpmc_b200host.so = the reference's host code
# linked against the GPU drop-in library instead of its own stiffness.c / solver.c / constitutive.c
REF_SO = Path(os.environ.get("LPMB_REF_SO", HERE / "_ref" / "liblpmc_ref.so"))

c_dp = C.POINTER(C.c_double)
c_dpp = C.POINTER(c_dp)
c_dppp = C.POINTER(c_dpp)
c_ip = C.POINTER(C.c_int)
c_ipp = C.POINTER(c_ip)


class DispBCPara(C.Structure):  # include/lpm.h:83-88
    _fields_ = [("type", C.c_int), ("flag", C.c_char), ("step", C.c_double)]


class ForceBCPara(C.Structure):  # include/lpm.h:90-99
    _fields_ = [("type", C.c_int), ("flag1", C.c_char), ("step1", C.c_double), ("flag2", C.c_char),
                ("step2", C.c_double), ("flag3", C.c_char), ("step3", C.c_double)]


def available() -> bool:
    return REF_SO.exists()


class RefLPM:
    """One per process (the reference keeps all state in process globals)."""

    _instance = None

    def __init__(self, so_path: os.PathLike | None = None):
        if RefLPM._instance is not None:
            raise RuntimeError("RefLPM is a process singleton (reference state is global)")
        self.lib = C.CDLL(str(so_path or REF_SO), mode=C.RTLD_GLOBAL)
        RefLPM._instance = self
        L = self.lib
        for fn in ("lpmb_ref_gather_d2", "lpmb_ref_scatter_d2", "lpmb_ref_gather_i2", "lpmb_ref_scatter_i2",
                   "lpmb_ref_gather_d3", "lpmb_ref_scatter_d3"):
            getattr(L, fn).restype = None
        L.allocInt1D.restype = c_ip
        L.allocInt1D.argtypes = [C.c_int, C.c_int]
        L.allocDouble1D.restype = c_dp
        L.allocDouble1D.argtypes = [C.c_int, C.c_double]
        L.allocDouble2D.restype = c_dpp
        L.allocDouble2D.argtypes = [C.c_int, C.c_int, C.c_double]
        L.setTypeRect.argtypes = [C.c_double] * 6 + [C.c_int]
        L.setTypeRect.restype = None
        L.cblas_dnrm2.restype = C.c_double
        L.cblas_dnrm2.argtypes = [C.c_int, c_dp, C.c_int]
        L.searchAFEMNeighbor.restype = C.c_int
        L.updateDamageGeneral.restype = C.c_int
        L.updateDamageGeneral.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.countNEqual.restype = C.c_int
        L.countNEqual.argtypes = [c_ip, C.c_int, C.c_int]
        L.lpmb_shim_set_threads.argtypes = [C.c_int]
        L.lpmb_shim_spmv_seconds.restype = C.c_double
        L.lpmb_shim_spmv_calls.restype = C.c_long
        L.omp_set_num_threads.argtypes = [C.c_int]

    @classmethod
    def instance(cls) -> "RefLPM":
        return cls._instance or cls()

    # ------------------------------------------------------------------ scalars
    def gi(self, name: str) -> int:
        return C.c_int.in_dll(self.lib, name).value

    def si(self, name: str, v: int) -> None:
        C.c_int.in_dll(self.lib, name).value = int(v)

    def gd(self, name: str) -> float:
        return C.c_double.in_dll(self.lib, name).value

    def sd(self, name: str, v: float) -> None:
        C.c_double.in_dll(self.lib, name).value = float(v)

    def darr(self, name: str, n: int):
        """fixed-size global double array (box[6], R_matrix[9], ...) as a ctypes view"""
        return (C.c_double * n).in_dll(self.lib, name)

    def iarr(self, name: str, n: int):
        return (C.c_int * n).in_dll(self.lib, name)

    def set_ptr(self, name: str, ptr) -> None:
        """point a global pointer variable (int*, double**, ...) at freshly allocated memory"""
        C.c_void_p.in_dll(self.lib, name).value = C.cast(ptr, C.c_void_p).value

    # ------------------------------------------------------------------ arrays
    @property
    def N(self) -> int:
        return self.gi("nparticle")

    @property
    def nn(self) -> int:
        return self.gi("nneighbors")

    @property
    def dim(self) -> int:
        return self.gi("dim")

    def d1(self, name: str, n: int) -> np.ndarray:
        p = c_dp.in_dll(self.lib, name)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def set_d1(self, name: str, a: np.ndarray) -> None:
        p = c_dp.in_dll(self.lib, name)
        a = np.ascontiguousarray(a, dtype=np.float64)
        C.memmove(p, a.ctypes.data, a.nbytes)

    def i1(self, name: str, n: int) -> np.ndarray:
        p = c_ip.in_dll(self.lib, name)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def set_i1(self, name: str, a: np.ndarray) -> None:
        p = c_ip.in_dll(self.lib, name)
        a = np.ascontiguousarray(a, dtype=np.int32)
        C.memmove(p, a.ctypes.data, a.nbytes)

    def d2(self, name: str, rows: int, cols: int) -> np.ndarray:
        out = np.empty((rows, cols), dtype=np.float64)
        self.lib.lpmb_ref_gather_d2(c_dpp.in_dll(self.lib, name), rows, cols, out.ctypes.data_as(c_dp))
        return out

    def set_d2(self, name: str, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.lib.lpmb_ref_scatter_d2(c_dpp.in_dll(self.lib, name), a.shape[0], a.shape[1], a.ctypes.data_as(c_dp))

    def i2(self, name: str, rows: int, cols: int) -> np.ndarray:
        out = np.empty((rows, cols), dtype=np.int32)
        self.lib.lpmb_ref_gather_i2(c_ipp.in_dll(self.lib, name), rows, cols, out.ctypes.data_as(c_ip))
        return out

    def set_i2(self, name: str, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.int32)
        self.lib.lpmb_ref_scatter_i2(c_ipp.in_dll(self.lib, name), a.shape[0], a.shape[1], a.ctypes.data_as(c_ip))

    def d3(self, name: str, rows: int, cols: int, depth: int) -> np.ndarray:
        out = np.empty((rows, cols, depth), dtype=np.float64)
        self.lib.lpmb_ref_gather_d3(c_dppp.in_dll(self.lib, name), rows, cols, depth, out.ctypes.data_as(c_dp))
        return out

    def set_d3(self, name: str, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.lib.lpmb_ref_scatter_d3(c_dppp.in_dll(self.lib, name), a.shape[0], a.shape[1], a.shape[2],
                                     a.ctypes.data_as(c_dp))

    # named groups (shapes from include/lpm.h + initialization.c:1120-1194)
    BOND_D2 = ("distance", "distance_initial", "csx", "csy", "csz", "csx_initial", "csy_initial", "csz_initial",
               "dL", "dL_ave", "ddL", "ddLp", "Kn", "Tv", "F", "F_temp", "bond_stress", "damage_broken", "damage_w")
    PART2_D2 = ("dL_total", "TdL_total", "ddL_total", "TddL_total", "damage_local", "damage_nonlocal")
    PART3_D2 = ("xyz", "xyz_initial", "xyz_temp", "J2_alpha", "J2_beta_eq")
    PART6_D2 = ("stress_tensor", "strain_tensor")
    PART_D1 = ("J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "sigmay", "damage_visual")
    DOF_D1 = ("residual", "Pex", "Pex_temp", "disp")
    BOND_I2 = ("neighbors", "nsign")
    PART_I1 = ("nb", "nb_initial", "nb_conn", "type", "pl_flag", "state_v")

    def get(self, name: str) -> np.ndarray:
        N, nn, dim = self.N, self.nn, self.dim
        if name in self.BOND_D2:
            return self.d2(name, N, nn)
        if name in self.PART2_D2:
            return self.d2(name, N, 2)
        if name in self.PART3_D2:
            return self.d2(name, N, 3)
        if name in self.PART6_D2:
            return self.d2(name, N, 6)
        if name in self.PART_D1:
            return self.d1(name, N)
        if name in self.DOF_D1:
            return self.d1(name, dim * N)
        if name == "Pin":
            return self.d1(name, 3 * N)
        if name in self.BOND_I2:
            return self.i2(name, N, nn)
        if name in self.PART_I1:
            return self.i1(name, N)
        if name in ("dispBC_index", "fix_index"):
            return self.i1(name, dim * N)
        if name == "conn":
            return self.i2(name, N, self.gi("nneighbors_AFEM") + 1)
        if name == "K_pointer":
            return self.i2(name, N + 1, 2)
        if name == "dLp":
            return self.d3(name, N, nn, 3)
        if name == "damage_D":
            return self.d3(name, N, nn, 2)
        if name == "J2_beta":
            return self.d3(name, N, 6, 3)
        if name in ("K_global", "JK"):
            nnz = int(self.i2("K_pointer", N + 1, 2)[N, 1])
            return self.d1(name, nnz) if name == "K_global" else self.i1(name, nnz)
        if name == "IK":
            return self.i1(name, dim * N + 1)
        if name == "Ce":
            return self.d2(name, self.gi("ntype"), 3)
        if name == "KnTve":
            return self.d2(name, self.gi("ntype"), 2 if self.gi("lattice") == 1 else 3)
        if name == "reaction_force":
            nbc = int(np.count_nonzero(self.i1("dispBC_index", dim * N) != 1))
            return self.d1(name, nbc)
        raise KeyError(name)

    def put(self, name: str, a: np.ndarray) -> None:
        a = np.asarray(a)
        if a.dtype.kind == "f":
            if a.ndim == 1:
                self.set_d1(name, a)
            elif a.ndim == 2:
                self.set_d2(name, a)
            else:
                self.set_d3(name, a)
        else:
            if a.ndim == 1:
                self.set_i1(name, a)
            else:
                self.set_i2(name, a)

    def snapshot(self, names) -> dict:
        return {n: self.get(n) for n in names}

    # ------------------------------------------------------------------ set-up
    def threads(self, nt: int) -> None:
        """OpenMP threads for the reference loops and the shim (1 = deterministic parity mode)."""
        self.lib.omp_set_num_threads(int(nt))
        self.lib.lpmb_shim_set_threads(int(nt))

    def setup_sc(self, box=(-0.2, 10.2, -0.2, 10.2, -0.2, 10.2), radius=0.2499999944120646, E0=146e3, mu0=0.3,
                 plmode=0, sigmay=200.0, J2_xi=0.0, J2_H=38.714e3, nbreak=20, critical_bstrain=1.0e-2,
                 damageb_A=10.0, damagec_A=0.0, damage_threshold=0.9, damage_L=0.5, dtime=0.01,
                 top_z=None, neighbor_search=True):
        """Re-play src/lpmc_project.c:75-345 (default 3-D simple-cubic cyclic case) with parameters.

        Types as in the default driver: 1 = top z-layer, 2 = bottom z-layer, 3 = full neighbour list
        (lpmc_project.c:179-182).  `top_z` defaults to box zmax - 0.2 (10.0 for the default box).
        """
        L = self.lib
        self.si("lattice", 2)
        self.si("dim", 3)
        self.sd("radius", radius)
        pbc = self.iarr("pbc", 3)
        pbc[0] = pbc[1] = pbc[2] = 0
        self.si("eulerflag", 0)
        PI = 3.14159265358979323846
        self.sd("angle1", PI / 180.0 * 0.0)
        self.sd("angle2", PI / 180.0 * 0.0)
        self.sd("angle3", PI / 180.0 * 0.0)
        b = self.darr("box", 6)
        for k in range(6):
            b[k] = box[k]
        self.sd("box_x", b[1] - b[0])
        self.sd("box_y", b[3] - b[2])
        self.sd("box_z", b[5] - b[4])
        L.createCuboid()
        mv = (C.c_double * 3)(-0.0, -0.0, -0.0)
        L.moveParticle(mv)
        L.initMatrices()
        N = self.N
        self.set_d2("xyz_initial", self.d2("xyz", N, 3))
        if neighbor_search == "lattice":
            self.inject_sc_topology()   # O(N), bit-identical to the two O(N^2) calls below (tests/test_oracle_ref.py)
        elif neighbor_search:
            L.searchNormalNeighbor()
            L.searchAFEMNeighbor()
        else:
            return  # caller fills the topology globals, then calls finish_setup_sc()
        self._finish_sc(E0, mu0, plmode, sigmay, J2_xi, J2_H, nbreak, critical_bstrain, damageb_A, damagec_A,
                        damage_threshold, damage_L, dtime, top_z)

    def inject_sc_topology(self):
        """What searchNormalNeighbor() + searchAFEMNeighbor() (neighbor.c:9-141) leave in the reference's globals, computed
        in O(N) with numpy for a FULL simple-cubic block (createCuboid, lattice 2, no carving) and written into the arrays
        initMatrices() allocated; K_global / IK / JK are allocated with the reference's own allocators (neighbor.c:132-134).
        The reference's O(N^2) search needs ~18 min at 10^6 particles; this makes S1 = 100^3 reachable for the CPU arm of
        bench.py.  Lists are in ascending-j order with the shells interleaved (neighbor.c:29,40), distances and unit
        vectors use the same expressions (sqrt(dx^2+dy^2+dz^2), (x_i - x_j)/dis), conn = sorted unique union (:56-112)."""
        L = self.lib
        N, nn = self.N, self.nn
        assert self.gi("lattice") == 2 and self.gi("dim") == 3 and nn == 18
        xyz = self.d2("xyz", N, 3)
        h = 2.0 * self.gd("radius")
        ijk = np.rint((xyz - xyz.min(axis=0)) / h).astype(np.int64)
        nx, ny, nz = (int(v) + 1 for v in ijk.max(axis=0))
        assert nx * ny * nz == N, "inject_sc_topology: not a full simple-cubic block"
        grid = np.full((nx + 4, ny + 4, nz + 4), -1, dtype=np.int64)      # 2-cell apron of -1 around the block
        grid[ijk[:, 0] + 2, ijk[:, 1] + 2, ijk[:, 2] + 2] = np.arange(N)
        assert np.abs(xyz - (xyz.min(axis=0) + h * ijk)).max() < 1e-6 * h
        o1 = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
        o2 = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1) if abs(a) + abs(b) + abs(c) == 2]
        look = lambda o: grid[ijk[:, 0] + 2 + o[0], ijk[:, 1] + 2 + o[1], ijk[:, 2] + 2 + o[2]]
        big = np.iinfo(np.int64).max
        cand = np.stack([look(o) for o in o1 + o2], axis=1)
        shell = np.array([0] * 6 + [1] * 12, dtype=np.int32)
        key = np.where(cand >= 0, cand, big)
        order = np.argsort(key, axis=1, kind="stable")
        nbr = np.take_along_axis(cand, order, axis=1)
        sgn = np.where(nbr >= 0, shell[order], 0).astype(np.int32)        # unused slots: initMatrices' fill value of nsign is kept below
        cnt = (nbr >= 0).sum(axis=1).astype(np.int32)
        # distances / unit vectors with the reference's expressions (neighbor.c:18,23-26)
        j = np.where(nbr >= 0, nbr, 0)
        d = xyz[j] - xyz[:, None, :]
        dis = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
        c1, c2 = 1.01 * self.gd("neighbor1_cutoff"), 1.01 * self.gd("neighbor2_cutoff")
        valid = nbr >= 0
        assert np.all((dis[valid & (sgn == 0)] < c1)) and np.all((dis[valid & (sgn == 1)] > c1) & (dis[valid & (sgn == 1)] < c2))
        with np.errstate(invalid="ignore", divide="ignore"):
            cs = (xyz[:, None, :] - xyz[j]) / dis[..., None]
        old = {n: self.d2(n, N, nn) for n in ("csx_initial", "csy_initial", "csz_initial", "distance_initial")}
        old_sgn, old_nbr = self.i2("nsign", N, nn), self.i2("neighbors", N, nn)
        for k, n in enumerate(("csx_initial", "csy_initial", "csz_initial")):
            self.set_d2(n, np.where(valid, cs[..., k], old[n]))
        self.set_d2("distance_initial", np.where(valid, dis, old["distance_initial"]))
        self.set_i2("nsign", np.where(valid, sgn, old_sgn))
        self.set_i2("neighbors", np.where(valid, nbr, old_nbr).astype(np.int32))
        for name, sh, width in (("neighbors1", 0, self.gi("nneighbors1")), ("neighbors2", 1, self.gi("nneighbors2"))):
            sel = valid & (sgn == sh)
            k2 = np.where(sel, nbr, big)
            part = np.sort(k2, axis=1)[:, :width]
            prev = self.i2(name, N, width)
            self.set_i2(name, np.where(part != big, part, prev).astype(np.int32))
        self.set_i1("nb", cnt)
        self.set_i1("nb_initial", cnt)
        del cand, key, order, d, dis, cs, old
        # conn: direct neighbours, 1st-of-1st, 2nd-of-2nd (reached through an EXISTING intermediate), self included
        nA = self.gi("nneighbors_AFEM") + 1
        offs = {}
        for group in (o1, o2):
            for a in group:
                for b in group:
                    offs.setdefault((a[0] + b[0], a[1] + b[1], a[2] + b[2]), []).append(a)
        cols = []
        for o, vias in sorted(offs.items()):
            tgt = look(o) if max(abs(v) for v in o) <= 2 else None
            ok = np.zeros(N, dtype=bool)
            for a in vias:
                ok |= look(a) >= 0
            cols.append(np.where(ok, tgt, -1))
        for o in o1 + o2:
            if o not in offs:
                cols.append(look(o))
        cm = np.stack(cols, axis=1)
        cm = np.sort(np.where(cm >= 0, cm, big), axis=1)
        assert cm.shape[1] <= nA or np.all(cm[:, nA:] == big)
        cm = cm[:, :nA]
        nbc = (cm != big).sum(axis=1).astype(np.int32)
        prev = self.i2("conn", N, nA)
        self.set_i2("conn", np.where(cm != big, cm, prev).astype(np.int32))
        self.set_i1("nb_conn", nbc)
        k0 = ((cm != big) & (cm >= np.arange(N)[:, None])).sum(axis=1).astype(np.int64)
        kp = np.zeros((N + 1, 2), dtype=np.int64)
        kp[:N, 0] = k0
        kp[1:, 1] = np.cumsum(9 * k0 - 3)
        assert kp[N, 1] < 2 ** 31, "nnz_upper exceeds the reference's 32-bit K_pointer"
        prevkp = self.i2("K_pointer", N + 1, 2)
        kp[N, 0] = prevkp[N, 0]
        self.set_i2("K_pointer", kp.astype(np.int32))
        nnz = int(kp[N, 1])
        L.allocInt1D.restype, L.allocDouble1D.restype = c_ip, c_dp
        self.set_ptr("JK", L.allocInt1D(nnz, -1))
        self.set_ptr("IK", L.allocInt1D(3 * N + 1, -1))
        self.set_ptr("K_global", L.allocDouble1D(nnz, -1.0))
        return nnz

    def setup_ct_geometry(self):
        """Geometry and topology of examples/CT_sc_ductile_nonlocal.c (:60-160; BASELINE config 5 as shipped): simple-cubic
        plate 50 x 48 x 10, radius 0.33, two stacked pre-cracks and the two loading holes -> 75 030 particles; then
        initMatrices + searchNormalNeighbor + searchAFEMNeighbor (O(N^2): ~25 s single-threaded).  No material set-up."""
        L = self.lib
        self.si("lattice", 2)
        self.si("dim", 3)
        radius = 0.33
        self.sd("radius", radius)
        pbc = self.iarr("pbc", 3)
        pbc[0] = pbc[1] = pbc[2] = 0
        self.si("eulerflag", 0)
        for a in ("angle1", "angle2", "angle3"):
            self.sd(a, 0.0)
        b = self.darr("box", 6)
        for k, v in enumerate((0.0, 50.0, 0.0, 48.0, 0.0, 10.0)):
            b[k] = v
        self.sd("box_x", b[1] - b[0])
        self.sd("box_y", b[3] - b[2])
        self.sd("box_z", b[5] - b[4])
        L.createCuboid()
        mv = (C.c_double * 3)(-0.0, -0.0, -0.0)
        L.moveParticle(mv)
        L.createCrack.argtypes = [C.c_double] * 4
        L.createCrack.restype = None
        ca1, ca2, w, ch = -10.0, 18.0, 1.0, 24.021
        L.createCrack(ca1, ca2, w, ch)
        L.createCrack(ca1, ca2 + 3 * radius, 0.5 * w, ch)
        L.removeCircle.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_char]
        L.removeCircle.restype = None
        L.removeCircle((C.c_double * 3)(9.9, 13.13, 0.0), 5.0, b"z")
        L.removeCircle((C.c_double * 3)(9.9, 34.91, 0.0), 5.0, b"z")
        L.initMatrices()
        N = self.N
        self.set_d2("xyz_initial", self.d2("xyz", N, 3))
        L.searchNormalNeighbor()
        L.searchAFEMNeighbor()

    def setup_2d(self, lattice=1, box=(0.0, 0.064, 0.0, 0.064, 0.0, 1.0), radius=3.2e-3, E0=210e3, mu0=0.3, nbreak=2,
                 critical_bstrain=2.7e-2, crack=None, crack_w=0.0):
        """Re-play the set-up of examples/shear_hex_brittle.c (:60-231; lattice 1 = hexagonal) or
        examples/3_point_bending_sq_brittle.c (lattice 0 = square) on a small box: 2-D, elastic (plmode 6) with brittle
        bond breaking.  Types as in the hex example: 1 = top y-layer, 2 = bottom y-layer, 3 = full neighbour list.
        crack = (a1, a2, h): createCrack(a1, a2, 0, h) before the matrices, defineCrack(a1, a2, h) after (:83-84, :224)."""
        L = self.lib
        self.si("lattice", lattice)
        self.si("dim", 2)
        self.sd("radius", radius)
        pbc = self.iarr("pbc", 3)
        pbc[0] = pbc[1] = pbc[2] = 0
        self.si("eulerflag", 0)
        for a in ("angle1", "angle2", "angle3"):
            self.sd(a, 0.0)
        b = self.darr("box", 6)
        for k in range(6):
            b[k] = box[k]
        self.sd("box_x", b[1] - b[0])
        self.sd("box_y", b[3] - b[2])
        self.sd("box_z", b[5] - b[4])
        L.createCuboid()
        if crack is not None:
            L.createCrack.argtypes = [C.c_double] * 4
            L.createCrack.restype = None
            L.createCrack(crack[0], crack[1], crack_w, crack[2])   # crack_w > 0 removes a notch of that half-width (bending example :121-122)
        L.initMatrices()
        N = self.N
        self.set_d2("xyz_initial", self.d2("xyz", N, 3))
        L.searchNormalNeighbor()
        L.searchAFEMNeighbor()
        ytop, ybot = float(self.d2("xyz", N, 3)[:, 1].max()), float(self.d2("xyz", N, 3)[:, 1].min())
        ntype = 0
        self.set_ptr("type", L.allocInt1D(N, ntype))
        ntype += 1
        L.setTypeRect(-100.0, 100.0, ytop - 1.0 * radius, 100.0, -100.0, 100.0, ntype)
        ntype += 1
        L.setTypeRect(-100.0, 100.0, -100.0, ybot + 1.0 * radius, -100.0, 100.0, ntype)
        ntype += 1
        L.setTypeFullNeighbor(C.c_int(ntype))
        ntype += 1
        self.si("ntype", ntype)
        C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0)
        C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0)
        C44 = E0 / 2.0 / (1.0 + mu0)
        self.set_ptr("Ce", L.allocDouble2D(ntype, 3, 0.0))
        self.set_d2("Ce", np.tile(np.array([C11, C12, C44]), (ntype, 1)))
        self.si("plmode", 6)
        self.si("nbreak", nbreak)
        self.sd("critical_bstrain", critical_bstrain)
        self.sd("damageb_A", 10.0)
        self.sd("damagec_A", 0.0)
        self.sd("damage_threshold", 0.9)
        self.sd("damage_L", 0.5)
        self.sd("dtime", 0.01)
        if crack is not None:
            L.defineCrack.argtypes = [C.c_double] * 3
            L.defineCrack.restype = None
            L.defineCrack(crack[0], crack[1], crack[2])
        L.calcKnTv()
        L.computedL()
        L.slipSysDefine3D()

    def setup_fcc(self, box=(0.0, 10.0, 0.0, 10.0, 0.0, 10.0), radius=0.3, C11=107.3e3, C12=60.8e3, C44=28.3e3,
                  cp_tau0=1.6, cp_taus=30.0, cp_h0=100.0, cp_p=4.0, cp_q=1.0, cp_eta=1000.0, cp_maxloop=10, dtime=0.1,
                  top_z=None, lattice=3):
        """Re-play the set-up of examples/FCC_Al_R0.3_001_tension.c (:62-229, :339-345 of the default driver for the
        call order): FCC lattice, Al elastic constants, crystal plasticity (plmode 1, 24 slip systems).  That example
        does not compile against the reference's current src/ (SURVEY section 2), its library functions do.
        Types: 1 top layer, 2 x-line, 3 y-line, 4 top fix point, 5 lower layer (:160-164)."""
        L = self.lib
        self.si("lattice", lattice)   # 3 = FCC; 4 = BCC (same call order, initialization.c:391-560,636-815)
        self.si("dim", 3)
        self.sd("radius", radius)
        pbc = self.iarr("pbc", 3)
        pbc[0] = pbc[1] = pbc[2] = 0
        self.si("eulerflag", 0)
        for a in ("angle1", "angle2", "angle3"):
            self.sd(a, 0.0)
        b = self.darr("box", 6)
        for k in range(6):
            b[k] = box[k]
        self.sd("box_x", b[1] - b[0])
        self.sd("box_y", b[3] - b[2])
        self.sd("box_z", b[5] - b[4])
        L.createCuboid()
        L.moveParticle((C.c_double * 3)(-0.0, -0.0, -0.0))
        L.initMatrices()
        N = self.N
        self.set_d2("xyz_initial", self.d2("xyz", N, 3))
        L.searchNormalNeighbor()
        L.searchAFEMNeighbor()
        if top_z is None:
            top_z = box[5]
        ntype = 0
        self.set_ptr("type", L.allocInt1D(N, ntype))
        ntype += 1
        r = radius
        for args in ((-100.0, 100.0, -100.0, 100.0, top_z - 1.2 * r, 100.0), (-100.0, 1.2 * r, -100.0, 100.0, top_z - 1.2 * r, 100.0),
                     (-100.0, 100.0, -100.0, 1.2 * r, top_z - 1.2 * r, 100.0), (-100.0, 1.2 * r, -100.0, 1.2 * r, top_z - 1.2 * r, 100.0),
                     (-100.0, 100.0, -100.0, 100.0, -100.0, 1.2 * r)):
            L.setTypeRect(*args, ntype)
            ntype += 1
        self.si("ntype", ntype)
        self.set_ptr("Ce", L.allocDouble2D(ntype, 3, 0.0))
        self.set_d2("Ce", np.tile(np.array([C11, C12, C44]), (ntype, 1)))
        self.si("plmode", 1)
        self.sd("cp_maxloop", cp_maxloop)
        t0, ts = self.darr("cp_tau0", 3), self.darr("cp_taus", 3)
        t0[0], t0[1], t0[2] = cp_tau0, 0.0, 0.0
        ts[0], ts[1], ts[2] = cp_taus, 0.0, 0.0
        self.sd("cp_h0", cp_h0)
        self.sd("cp_p", cp_p)
        self.sd("cp_q", cp_q)
        self.sd("cp_eta", cp_eta)
        self.sd("dtime", dtime)
        self.si("nbreak", 20)
        self.sd("critical_bstrain", 1.0e-2)
        self.sd("damage_threshold", 0.9)
        self.sd("damage_L", 0.5)
        L.calcKnTv()
        L.computedL()
        L.slipSysDefine3D()
        L.computeCab()

    # crystal-plasticity arrays (allocated by slipSysDefine3D, initialization.c:570,817-826)
    def get_cp(self, name: str) -> np.ndarray:
        N, S = self.N, self.gi("nslipSys")
        if name in ("cp_gy", "cp_A_single"):
            return self.d3(name, N, S, 3)
        if name == "cp_A":
            return self.d2(name, N, 3)
        if name == "cp_Cab":
            return self.d2(name, N, S * S)
        if name in ("cp_RSS", "cp_dgy", "cp_dA_single"):
            return self.d2(name, N, S)
        if name == "cp_Jact":
            return self.i2(name, N, S)
        if name == "cp_dA":
            return self.d1(name, N)
        if name == "schmid_tensor":
            return self.d2(name, S, 6)
        raise KeyError(name)

    def _finish_sc(self, E0, mu0, plmode, sigmay, J2_xi, J2_H, nbreak, critical_bstrain, damageb_A, damagec_A,
                   damage_threshold, damage_L, dtime, top_z):
        L = self.lib
        N = self.N
        radius = self.gd("radius")
        if top_z is None:
            top_z = self.darr("box", 6)[5] - 0.2
        elif top_z == "auto":  # other boxes: the actual top lattice layer
            top_z = float(self.d2("xyz", N, 3)[:, 2].max())
        ntype = 0
        self.set_ptr("type", L.allocInt1D(N, ntype))
        ntype += 1
        L.setTypeRect(-100.0, 100.0, -100.0, 100.0, top_z - 1.2 * radius, 100.0, ntype)
        ntype += 1
        L.setTypeRect(-100.0, 100.0, -100.0, 100.0, -100.0, 1.2 * radius, ntype)
        ntype += 1
        L.setTypeFullNeighbor(C.c_int(ntype))
        ntype += 1
        self.si("ntype", ntype)
        C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0)
        C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0)
        C44 = E0 / 2.0 / (1.0 + mu0)
        self.set_ptr("Ce", L.allocDouble2D(ntype, 3, 0.0))
        self.set_d2("Ce", np.tile(np.array([C11, C12, C44]), (ntype, 1)))
        self.si("plmode", plmode)
        self.set_ptr("sigmay", L.allocDouble1D(N, sigmay))
        self.sd("J2_xi", J2_xi)
        self.sd("J2_H", J2_H)
        self.si("nbreak", nbreak)
        self.sd("critical_bstrain", critical_bstrain)
        self.sd("damageb_A", damageb_A)
        self.sd("damagec_A", damagec_A)
        self.sd("damage_threshold", damage_threshold)
        self.sd("damage_L", damage_L)
        self.sd("dtime", dtime)
        L.calcKnTv()
        L.computedL()
        L.slipSysDefine3D()
        if self.gi("lattice") in (3, 4):
            L.computeCab()

    # ------------------------------------------------------------------ driver loop
    def norms(self):
        dim, N = self.dim, self.N
        nr = self.lib.cblas_dnrm2(dim * N, c_dp.in_dll(self.lib, "residual"), 1)
        nbc = self.lib.countNEqual(c_ip.in_dll(self.lib, "dispBC_index"), N * dim, 1)
        nf = self.lib.cblas_dnrm2(nbc, c_dp.in_dll(self.lib, "reaction_force"), 1)
        return nr, nf

    def begin_step(self, dbp, fbp, load_indicator=1, hooks=None):
        """src/lpmc_project.c:387-414: save temps, FD tangent, BCs, predictor, residual."""
        L = self.lib
        hooks = hooks or {}
        N, nn, dim = self.N, self.nn, self.dim
        self.set_d2("xyz_temp", self.d2("xyz", N, 3))
        self.set_d2("F_temp", self.d2("F", N, nn))
        self.set_d1("Pex_temp", self.d1("Pex", dim * N))
        hooks.get("stiffness", lambda: (L.calcStiffness2DFiniteDifference(6) if dim == 2
                                        else L.calcStiffness3DFiniteDifference(6)))()
        d_arr = (DispBCPara * max(1, len(dbp)))(*[DispBCPara(t, f.encode(), s) for (t, f, s) in dbp])
        f_arr = (ForceBCPara * max(1, len(fbp)))(*[ForceBCPara(t, b"x", sx, b"y", sy, b"z", sz)
                                                    for (t, sx, sy, sz) in fbp])
        L.setDispBC(len(dbp), d_arr)
        L.setForceBC(len(fbp), f_arr)
        hooks.get("predictor", lambda: L.computeBondForceGeneral(4, load_indicator))()
        L.updateRR()
        return self.norms()

    def newton_iteration(self, load_indicator=1, hooks=None):
        """One pass of src/lpmc_project.c:426-464.  Returns the new residual norm."""
        L = self.lib
        hooks = hooks or {}
        dim = self.dim
        L.switchStateV(0)
        if dim == 2:
            L.setDispBC_stiffnessUpdate2D()
        else:
            L.setDispBC_stiffnessUpdate3D()
        hooks.get("solve", L.solverCG)()
        hooks.get("bondforce", lambda: L.computeBondForceGeneral(self.gi("plmode"), load_indicator))()
        L.updateRR()
        return self.norms()[0]

    def load_step(self, step_no, dbp, fbp, load_indicator=1, hooks=None, bond_file=b"/dev/null", max_iter=100,
                  on_iter=None):
        """One load step of the default driver loop (src/lpmc_project.c:382-546), without writers.
        Returns (n_newton_iterations, broken_bonds_total)."""
        L = self.lib
        TOLITER = 1e-4
        nr, nf = self.begin_step(dbp, fbp, load_indicator, hooks)
        total_ni, total_broken = 0, 0
        while True:
            tol_mult = max(nr, nf)
            ni = 0
            while nr > TOLITER * tol_mult and ni < max_iter:
                nr = self.newton_iteration(load_indicator, hooks)
                ni += 1
                if on_iter:
                    on_iter(step_no, ni, nr)
            total_ni += ni
            L.computeStrain()
            broken = L.updateDamageGeneral(bond_file, step_no, self.gi("plmode"))
            L.updateCrack()
            L.switchStateV(1)
            total_broken += broken
            if broken <= 0:
                break
            (hooks or {}).get("stiffness", lambda: (L.calcStiffness2DFiniteDifference(6) if self.dim == 2
                                                    else L.calcStiffness3DFiniteDifference(6)))()
            L.updateRR()
            nr, nf = self.norms()
        return total_ni, total_broken
