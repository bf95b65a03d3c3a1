"""ctypes binding of include/lpmb200.h (liblpmb200.so).  Host-side plumbing only.

Every method forwards to one C-ABI entry point; errors raise LPMBError carrying
lpmb_last_error().  There is deliberately no fallback: without the built library the import
fails, and without a GPU `Context(...)` raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
lib_path = _HERE / "liblpmb200.so"
HEADER = _HERE.parent / "include" / "lpmb200.h"


class LPMBError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lpmb error {code}: {msg}")
        self.code = code


if not lib_path.exists():
    raise ImportError(f"{lib_path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). The CUDA library is the product; there is no fallback.")
lib = C.CDLL(str(lib_path))

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_vp = C.c_void_p

lib.lpmb_last_error.restype = C.c_char_p
lib.lpmb_launch_count.restype = C.c_longlong
lib.lpmb_launch_count.argtypes = [c_vp]
lib.lpmb_stream.restype = c_vp
lib.lpmb_stream.argtypes = [c_vp]
lib.lpmb_spmv_bytes.restype = C.c_longlong
lib.lpmb_spmv_bytes.argtypes = [c_vp]
lib.lpmb_spmv_bytes_stored.restype = C.c_longlong
lib.lpmb_spmv_bytes_stored.argtypes = [c_vp]
lib.lpmb_spmv_bytes_bricks.restype = C.c_longlong
lib.lpmb_spmv_bytes_bricks.argtypes = [c_vp]
lib.lpmb_matrix_enable_bricks.argtypes = [c_vp, C.c_int]
lib.lpmb_dist_mode.argtypes = [c_vp]
lib.lpmb_compute_strain.argtypes = [c_vp]
lib.lpmb_destroy.restype = None
lib.lpmb_destroy.argtypes = [c_vp]
lib.lpmb_create.argtypes = [C.POINTER(c_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
lib.lpmb_set_param.argtypes = [c_vp, C.c_char_p, C.c_double]
lib.lpmb_get_param.argtypes = [c_vp, C.c_char_p, c_dp]
lib.lpmb_field_set.argtypes = [c_vp, C.c_char_p, c_vp, C.c_size_t]
lib.lpmb_field_get.argtypes = [c_vp, C.c_char_p, c_vp, C.c_size_t]
lib.lpmb_field_device.argtypes = [c_vp, C.c_char_p, C.POINTER(c_vp), C.POINTER(C.c_size_t)]
lib.lpmb_fields_get_staged.argtypes = [c_vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_vp), C.POINTER(C.c_size_t)]
lib.lpmb_set_neighbors.argtypes = [c_vp, c_vp, c_vp]
lib.lpmb_set_connectivity.argtypes = [c_vp, c_vp]
lib.lpmb_build_topology.argtypes = [c_vp, C.c_double, C.c_double]
lib.lpmb_csr_sizes.argtypes = [c_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
lib.lpmb_get_k_pointer.argtypes = [c_vp, c_vp]
lib.lpmb_matrix_from_upper_csr.argtypes = [c_vp, c_vp, C.c_longlong]
lib.lpmb_matrix_to_upper_csr.argtypes = [c_vp, c_vp, c_vp, c_vp]
lib.lpmb_fd_stiffness.argtypes = [c_vp, C.c_int]
lib.lpmb_matrix_fill_test_pattern.argtypes = [c_vp]
lib.lpmb_spmv_host.argtypes = [c_vp, c_vp, c_vp]
lib.lpmb_spmv_bench.argtypes = [c_vp, C.c_int, C.c_int, c_dp]
lib.lpmb_set_dof_mask.argtypes = [c_vp, c_vp, c_vp]
lib.lpmb_solve_cg.argtypes = [c_vp, c_vp, c_vp, C.c_double, C.c_double, C.c_int, C.c_int, c_ip]
lib.lpmb_solve_cg_device.argtypes = [c_vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_ip]
lib.lpmb_calc_kntv.argtypes = [c_vp, c_vp, C.c_int]
lib.lpmb_compute_dl.argtypes = [c_vp]
lib.lpmb_bond_force.argtypes = [c_vp, C.c_int, C.c_int]
lib.lpmb_bond_force_particle.argtypes = [c_vp, C.c_int, C.c_int, C.c_int]
lib.lpmb_snapshot_save.argtypes = [c_vp, C.c_char_p]
lib.lpmb_snapshot_load.argtypes = [c_vp, C.c_char_p]
lib.lpmb_switch_state.argtypes = [c_vp, C.c_int]
lib.lpmb_update_rr.argtypes = [c_vp, c_dp, c_dp]
lib.lpmb_update_damage.argtypes = [c_vp, C.c_int, c_ip, c_vp, C.c_int]
lib.lpmb_update_crack.argtypes = [c_vp]
lib.lpmb_newton_iteration.argtypes = [c_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, c_ip, c_dp]
lib.lpmb_dist_unique_id.argtypes = [c_vp]
lib.lpmb_dist_init.argtypes = [c_vp, c_vp, C.c_int, C.c_int]
lib.lpmb_dist_set_slab.argtypes = [c_vp] + [C.c_int] * 8
lib.lpmb_dist_exchange_field.argtypes = [c_vp, C.c_char_p, C.c_int]
lib.lpmb_brittle_select.argtypes = [C.c_int, c_vp, c_vp, C.c_int, c_vp]
lib.lpmb_mg_slab_plan.argtypes = [C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
lib.lpmb_synchronize.argtypes = [c_vp]
lib.lpmb_set_schmid_tensor.argtypes = [c_vp, c_vp, C.c_int]
lib.lpmb_compute_cab.argtypes = [c_vp]
lib.lpmb_set_profiling.argtypes = [c_vp, C.c_int]
lib.lpmb_get_profile.argtypes = [c_vp, c_dp, C.POINTER(C.c_longlong)]
lib.lpmb_field_copy.argtypes = [c_vp, C.c_char_p, C.c_char_p]
lib.lpmb_apply_disp_bc.argtypes = [c_vp, C.c_int, C.c_char, C.c_double]
lib.lpmb_apply_force_bc.argtypes = [c_vp, C.c_int, C.c_double, C.c_double, C.c_double]


def declared_symbols() -> list[str]:
    """every function name declared in include/lpmb200.h (used by the CPU symbol test)"""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lpmb_[a-z0-9_]+)\s*\(", text)))


def brittle_select(keys, strains, nbreak: int):
    """lpmb_brittle_select: (keys, strains) of the bonds that break, in the reference's order (host arithmetic, no GPU)"""
    keys = np.ascontiguousarray(keys, dtype=np.int64).copy()
    strains = np.ascontiguousarray(strains, dtype=np.float64).copy()
    first = C.c_int()
    _check(lib.lpmb_brittle_select(len(keys), keys.ctypes.data, strains.ctypes.data, int(nbreak), C.addressof(first)))
    return keys[first.value:], strains[first.value:]


def mg_slab_plan(world: int, rank: int, owned, nx: int, ny: int, nz_local0: int, ghost_lo0: int, max_levels: int = 12):
    """levels of the fast mode's multigrid hierarchy and rank `rank`'s part of each (lpmb_mg_slab_plan: host arithmetic, no GPU)"""
    owned = np.ascontiguousarray(owned, dtype=np.int64)
    plan = np.zeros((max_levels, 8), dtype=np.int32)
    off, cnt = np.zeros(16, dtype=np.int64), np.zeros(16, dtype=np.int64)
    nlev, lrep = C.c_int(), C.c_int()
    _check(lib.lpmb_mg_slab_plan(world, rank, owned.ctypes.data, nx, ny, nz_local0, ghost_lo0, max_levels, plan.ctypes.data, off.ctypes.data,
                                 cnt.ctypes.data, C.addressof(nlev), C.addressof(lrep)))
    keys = ("dist", "nx", "ny", "nz", "gz0", "oz0", "oz1", "nzg")
    return {"levels": [dict(zip(keys, (int(v) for v in plan[l]))) for l in range(nlev.value)], "lrep": lrep.value,
            "gat_off": off[:world].tolist(), "gat_cnt": cnt[:world].tolist()}


def device_count() -> int:
    return int(lib.lpmb_device_count())


def _check(rc: int, ok=(0,)):
    if rc not in ok:
        raise LPMBError(rc, lib.lpmb_last_error().decode())
    return rc


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# host dtype of the integer fields (everything else is float64)
_INT_FIELDS = {"cp_Jact", "neighbors", "nsign", "mirror", "oppslot", "type", "pl_flag", "nb", "nb_initial", "dispBC_index",
               "fix_index", "state_v"}

NOTCONVERGED = 4
# lpmb_update_damage codes of the two laws updateDamageGeneral's dispatcher keeps commented out (include/lpmb200.h)
DAMAGE_PWISE_LOCAL, DAMAGE_BWISE_NONLOCAL = 100, 101


class Context:
    """One device context = one particle system (or one slab of it) resident in HBM."""

    def __init__(self, nparticle: int, dim: int, lattice: int, nneighbors: int, nconn_max: int, device: int = 0):
        self._h = c_vp()
        _check(lib.lpmb_create(C.byref(self._h), device, nparticle, dim, lattice, nneighbors, nconn_max))
        self.N, self.dim, self.lattice, self.nn, self.nconn = nparticle, dim, lattice, nneighbors, nconn_max

    def close(self):
        if self._h:
            lib.lpmb_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters / fields
    def set_param(self, name: str, v: float):
        _check(lib.lpmb_set_param(self._h, name.encode(), float(v)))

    def set_params(self, **kw):
        for k, v in kw.items():
            self.set_param(k, v)

    def set_field(self, name: str, arr):
        a = _i32(arr) if name in _INT_FIELDS else _f64(arr)
        _check(lib.lpmb_field_set(self._h, name.encode(), a.ctypes.data, a.size))

    def _shape(self, name: str):
        N, nn, dim = self.N, self.nn, self.dim
        comps = {"xyz": 3, "xyz_initial": 3, "xyz_temp": 3, "xyz_save": 3, "dL_total": 2, "TdL_total": 2, "ddL_total": 2,
                 "TddL_total": 2, "stress_tensor": 6, "strain_tensor": 6, "J2_beta0": 6, "J2_beta1": 6, "J2_beta2": 6}
        if name in comps:
            return (N, comps[name])
        S = getattr(self, "nslip", 0)
        if name == "cp_Cab":
            return (N, S * S)
        if name in ("cp_RSS", "cp_Jact", "cp_dgy", "cp_dA_single") or name[:-1] in ("cp_gy", "cp_A_single"):
            return (N, S)
        if name in ("residual", "residual_save", "Pex", "Pex_temp", "disp", "dispBC_index", "fix_index"):
            return (N * dim,)
        if name == "Pin":
            return (N * 3,)
        bond = {"distance", "distance_initial", "csx", "csy", "csz", "csx_initial", "csy_initial", "csz_initial", "dL",
                "dL_ave", "ddL", "ddLp", "Kn", "Tv", "F", "F_temp", "bond_stress", "damage_broken", "damage_w", "dLp0",
                "dLp1", "dLp2", "damage_D0", "damage_D1", "neighbors", "nsign", "mirror", "oppslot"}
        if name in bond:
            return (N, nn)
        return (N,)

    def get_field(self, name: str) -> np.ndarray:
        out = np.empty(self._shape(name), dtype=np.int32 if name in _INT_FIELDS else np.float64)
        _check(lib.lpmb_field_get(self._h, name.encode(), out.ctypes.data, out.size))
        return out

    def get_fields(self, names) -> dict:
        """several fields with one device->host copy and one synchronisation (lpmb_fields_get_staged)"""
        n = len(names)
        arr = (C.c_char_p * n)(*[nm.encode() for nm in names])
        ptrs, cnts = (c_vp * n)(), (C.c_size_t * n)()
        _check(lib.lpmb_fields_get_staged(self._h, n, arr, ptrs, cnts))
        out = {}
        for k, nm in enumerate(names):
            dt = np.int32 if nm in _INT_FIELDS else np.float64
            buf = (C.c_char * (cnts[k] * np.dtype(dt).itemsize)).from_address(ptrs[k])
            out[nm] = np.frombuffer(buf, dtype=dt).reshape(self._shape(nm)).copy()
        return out

    # -- topology
    def set_neighbors(self, neighbors, nsign):
        a, b = _i32(neighbors), _i32(nsign)
        assert a.shape == (self.N, self.nn) and b.shape == a.shape
        _check(lib.lpmb_set_neighbors(self._h, a.ctypes.data, b.ctypes.data))

    def set_connectivity(self, conn):
        a = _i32(conn)
        assert a.shape == (self.N, self.nconn), (a.shape, self.N, self.nconn)
        _check(lib.lpmb_set_connectivity(self._h, a.ctypes.data))

    def build_topology(self, cutoff1: float, cutoff2: float):
        _check(lib.lpmb_build_topology(self._h, cutoff1, cutoff2))

    def csr_sizes(self):
        nnz, nblk = C.c_longlong(), C.c_longlong()
        _check(lib.lpmb_csr_sizes(self._h, C.byref(nnz), C.byref(nblk)))
        return nnz.value, nblk.value

    def k_pointer(self) -> np.ndarray:
        out = np.empty((self.N + 1, 2), dtype=np.int32)
        _check(lib.lpmb_get_k_pointer(self._h, out.ctypes.data))
        return out

    # -- matrix
    def matrix_from_upper_csr(self, K_global):
        a = _f64(K_global)
        _check(lib.lpmb_matrix_from_upper_csr(self._h, a.ctypes.data, a.size))

    def matrix_to_upper_csr(self):
        nnz, _ = self.csr_sizes()
        K = np.empty(nnz, dtype=np.float64)
        JK = np.empty(nnz, dtype=np.int32)
        IK = np.empty(self.N * self.dim + 1, dtype=np.int32)
        _check(lib.lpmb_matrix_to_upper_csr(self._h, K.ctypes.data, IK.ctypes.data, JK.ctypes.data))
        return K, IK, JK

    def fd_stiffness(self, emulate_side_effects: bool = False):
        _check(lib.lpmb_fd_stiffness(self._h, int(emulate_side_effects)))

    def fill_test_pattern(self):
        _check(lib.lpmb_matrix_fill_test_pattern(self._h))

    def spmv(self, x) -> np.ndarray:
        x = _f64(x)
        y = np.empty_like(x)
        _check(lib.lpmb_spmv_host(self._h, x.ctypes.data, y.ctypes.data))
        return y

    def spmv_bench(self, reps: int = 20, variant: int = 0) -> float:
        ms = C.c_double()
        _check(lib.lpmb_spmv_bench(self._h, reps, variant, C.byref(ms)))
        return ms.value

    def spmv_bytes(self) -> int:
        return int(lib.lpmb_spmv_bytes(self._h))

    def spmv_bytes_stored(self) -> int:
        return int(lib.lpmb_spmv_bytes_stored(self._h))

    def dist_mode(self) -> int:
        """0 single GPU, 1 NCCL only, 2 peer-memory scalars, 3 peer-memory scalars + halo push (include/lpmb200.h)"""
        return int(lib.lpmb_dist_mode(self._h))

    def enable_bricks(self, on: bool = True):
        """brick-blocked symmetric SpMV for the CG (include/lpmb200.h); raises if the lattice is not eligible"""
        _check(lib.lpmb_matrix_enable_bricks(self._h, 1 if on else 0))

    def spmv_bytes_bricks(self) -> int:
        return int(lib.lpmb_spmv_bytes_bricks(self._h))

    # -- solve
    def set_dof_mask(self, dispBC_index=None, fix_index=None):
        a = _i32(dispBC_index) if dispBC_index is not None else None
        b = _i32(fix_index) if fix_index is not None else None
        _check(lib.lpmb_set_dof_mask(self._h, a.ctypes.data if a is not None else None,
                                     b.ctypes.data if b is not None else None))

    def solve_cg(self, rhs, rel=1e-8, abs_tol=1e-12, maxit=None, use_mask=False):
        b = _f64(rhs)
        x = np.empty_like(b)
        it = C.c_int()
        rc = _check(lib.lpmb_solve_cg(self._h, b.ctypes.data, x.ctypes.data, rel, abs_tol,
                                      int(maxit or b.size), int(use_mask), C.byref(it)), ok=(0, NOTCONVERGED))
        return x, it.value, rc == 0

    def solve_cg_device(self, rel=1e-8, abs_tol=1e-12, maxit=None, use_mask=True, update_xyz=True):
        it = C.c_int()
        rc = _check(lib.lpmb_solve_cg_device(self._h, rel, abs_tol, int(maxit or self.N * self.dim), int(use_mask),
                                             int(update_xyz), C.byref(it)), ok=(0, NOTCONVERGED))
        return it.value, rc == 0

    # -- constitutive path
    def calc_kntv(self, Ce):
        a = _f64(Ce)
        _check(lib.lpmb_calc_kntv(self._h, a.ctypes.data, a.shape[0]))

    def compute_dl(self):
        _check(lib.lpmb_compute_dl(self._h))

    def bond_force(self, plmode: int, load_indicator: int = 1):
        _check(lib.lpmb_bond_force(self._h, plmode, load_indicator))

    def bond_force_particle(self, plmode: int, particle: int, load_indicator: int = 1):
        """the reference's per-particle law entry points: plmode 6 / 4 / 0 / 3 / 5 / 1 = computeBondForceElastic(ii),
        ...IncrementalUpdating(ii), ...J2mixedLinear3D(ii), ...J2energyReturnMap(ii, t), ...J2nonlinearIso(ii), ...CPMiehe(ii)"""
        _check(lib.lpmb_bond_force_particle(self._h, plmode, int(particle), load_indicator))

    def compute_strain(self):
        _check(lib.lpmb_compute_strain(self._h))

    def switch_state(self, flag: int):
        _check(lib.lpmb_switch_state(self._h, flag))

    def update_rr(self):
        a, b = C.c_double(), C.c_double()
        _check(lib.lpmb_update_rr(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def update_damage(self, plmode: int, max_pairs: int = 4096):
        broken = C.c_int()
        pairs = np.full((max_pairs, 2), -1, dtype=np.int32)
        _check(lib.lpmb_update_damage(self._h, plmode, C.byref(broken), pairs.ctypes.data, max_pairs))
        return broken.value, pairs[: min(broken.value, max_pairs)]

    def update_crack(self):
        _check(lib.lpmb_update_crack(self._h))

    def snapshot_save(self, path):
        _check(lib.lpmb_snapshot_save(self._h, str(path).encode()))

    def snapshot_load(self, path):
        _check(lib.lpmb_snapshot_load(self._h, str(path).encode()))

    def newton_iteration(self, plmode: int, load_indicator: int = 1, rel=1e-8, abs_tol=1e-12, maxit=None):
        it, nr = C.c_int(), C.c_double()
        _check(lib.lpmb_newton_iteration(self._h, plmode, load_indicator, rel, abs_tol,
                                         int(maxit or self.N * self.dim), C.byref(it), C.byref(nr)),
               ok=(0, NOTCONVERGED))
        return it.value, nr.value

    # -- crystal plasticity
    def set_schmid_tensor(self, schmid):
        a = _f64(schmid)
        self.nslip = a.shape[0]
        _check(lib.lpmb_set_schmid_tensor(self._h, a.ctypes.data, a.shape[0]))

    def compute_cab(self):
        _check(lib.lpmb_compute_cab(self._h))

    # -- multi-GPU
    @staticmethod
    def dist_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib.lpmb_dist_unique_id(buf))
        return buf.raw

    def dist_init(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == 128
        _check(lib.lpmb_dist_init(self._h, unique_id, rank, world))

    def dist_set_slab(self, own0, own1, nrl, nrh, nsl, nsh, wsl, wsh):
        _check(lib.lpmb_dist_set_slab(self._h, own0, own1, nrl, nrh, nsl, nsh, wsl, wsh))

    def dist_exchange_field(self, name: str, wide: bool = True):
        _check(lib.lpmb_dist_exchange_field(self._h, name.encode(), int(wide)))

    def set_profiling(self, on: bool = True):
        _check(lib.lpmb_set_profiling(self._h, int(on)))

    def get_profile(self):
        ms, calls = C.c_double(), C.c_longlong()
        _check(lib.lpmb_get_profile(self._h, C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

    def copy_field(self, dst: str, src: str):
        _check(lib.lpmb_field_copy(self._h, dst.encode(), src.encode()))

    def apply_disp_bc(self, type_: int, axis: str, step: float):
        _check(lib.lpmb_apply_disp_bc(self._h, type_, axis.encode(), step))

    def apply_force_bc(self, type_: int, sx: float, sy: float, sz: float):
        _check(lib.lpmb_apply_force_bc(self._h, type_, sx, sy, sz))

    def synchronize(self):
        _check(lib.lpmb_synchronize(self._h))

    @property
    def launches(self) -> int:
        return int(lib.lpmb_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(lib.lpmb_stream(self._h) or 0)
