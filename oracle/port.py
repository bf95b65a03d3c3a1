"""oracle/port.py -- ctypes front end of the flat-array C restatement (oracle/lpm_oracle.c).

TEST INFRASTRUCTURE ONLY (see the header of lpm_oracle.c).  `Port` keeps the state of one particle system in
numpy arrays laid out like the reference's logical arrays and exposes the hot-path steps by the reference's
function names."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

SO = Path(__file__).resolve().parent / "liblpm_oracle.so"


def available() -> bool:
    return SO.exists()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Port:
    def __init__(self, xyz, dim=3, nn=18, nconn=61, radius=0.25, particle_volume=None):
        self.lib = C.CDLL(str(SO))
        self.lib.oracle_cg.restype = C.c_int
        self.N, self.dim, self.nn, self.nconn = len(xyz), dim, nn, nconn
        self.radius = float(radius)
        self.V = float(particle_volume if particle_volume is not None else (2 * radius) ** 3)
        N = self.N
        f2 = lambda c: np.zeros((N, c))
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).copy()
        self.xyz_initial = self.xyz.copy()
        self.xyz_temp = self.xyz.copy()
        self.neighbors = np.full((N, nn), -1, np.int32)
        self.nsign = np.full((N, nn), -1, np.int32)
        self.nb = np.zeros(N, np.int32)
        self.nb_initial = np.zeros(N, np.int32)
        for n in ("distance_initial", "csx_initial", "csy_initial", "csz_initial", "distance", "dL", "dL_ave", "ddL", "ddLp", "csx", "csy",
                  "csz", "Kn", "Tv", "F", "F_temp", "bond_stress", "dLp0", "dLp1", "dLp2", "damage_D0", "damage_D1"):
            setattr(self, n, f2(nn))
        self.damage_broken, self.damage_w = np.ones((N, nn)), np.ones((N, nn))
        for n in ("dL_total", "TdL_total", "ddL_total", "TddL_total"):
            setattr(self, n, f2(2))
        self.stress_tensor = f2(6)
        self.J2_beta = [f2(6) for _ in range(3)]
        self.J2_alpha = [np.zeros(N) for _ in range(3)]
        for n in ("J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "sigmay", "damage_visual", "damage_nonlocal0"):
            setattr(self, n, np.zeros(N))
        self.type = np.zeros(N, np.int32)
        self.pl_flag = np.zeros(N, np.int32)
        self.Pin = np.zeros(3 * N)
        self.Pex = np.zeros(dim * N)
        self.residual = np.zeros(dim * N)
        self.disp = np.zeros(dim * N)
        self.dispBC_index = np.ones(dim * N, np.int32)
        self.fix_index = np.ones(dim * N, np.int32)
        self.J2_H = self.J2_xi = 0.0
        self.Ce = np.zeros((1, 3))

    # neighbor.c:9-141
    def search_neighbors(self, cutoff1, cutoff2):
        ov = self.lib.oracle_search_neighbors(self.N, _p(self.xyz), C.c_double(cutoff1), C.c_double(cutoff2), self.nn, _p(self.neighbors),
                                              _p(self.nsign), _p(self.nb), _p(self.distance_initial), _p(self.csx_initial),
                                              _p(self.csy_initial), _p(self.csz_initial))
        assert ov == 0
        self.nb_initial[:] = self.nb
        self.conn = np.full((self.N, self.nconn), -1, np.int32)
        self.nb_conn = np.zeros(self.N, np.int32)
        self.kp0 = np.zeros(self.N + 1, np.int64)
        self.kp1 = np.zeros(self.N + 1, np.int64)
        ov = self.lib.oracle_afem_conn(self.N, self.nn, self.nconn, self.dim, _p(self.neighbors), _p(self.nsign), _p(self.nb), _p(self.conn),
                                       _p(self.nb_conn), _p(self.kp0), _p(self.kp1))
        assert ov == 0
        nnz = int(self.kp1[-1])
        self.K_global, self.JK, self.IK = np.zeros(nnz), np.zeros(nnz, np.int32), np.zeros(self.dim * self.N + 1, np.int32)

    def _geometry(self, dLp, apply_broken=1, distance=None):
        self.lib.oracle_geometry(self.N, self.nn, _p(self.xyz), _p(self.neighbors), _p(self.nsign), _p(self.nb_initial),
                                 _p(self.distance_initial), _p(dLp), _p(self.damage_broken), _p(self.Tv), apply_broken, _p(self.dL),
                                 _p(self.csx), _p(self.csy), _p(self.csz), _p(self.dL_total), _p(self.TdL_total), _p(distance))

    def computedL(self):
        self._geometry(self.dLp0, 0, self.distance)

    def computeStress(self):
        self.lib.oracle_stress(self.N, self.nn, C.c_double(self.V), _p(self.nb), _p(self.nb_initial), _p(self.distance_initial),
                               _p(self.csx_initial), _p(self.csy_initial), _p(self.csz_initial), _p(self.damage_broken), _p(self.F),
                               _p(self.csx), _p(self.csy), _p(self.csz), _p(self.stress_tensor), _p(self.J2_stresseq), _p(self.J2_stressm),
                               _p(self.J2_triaxiality), _p(self.bond_stress))

    def switchStateV(self, flag):
        if flag == 0:
            self.dLp0[:] = self.dLp1; self.damage_D0[:] = self.damage_D1
            self.J2_beta[0][:] = self.J2_beta[1]; self.J2_alpha[0][:] = self.J2_alpha[1]
        elif flag == 1:
            self.dLp1[:] = self.dLp0; self.damage_D1[:] = self.damage_D0
            self.J2_beta[1][:] = self.J2_beta[0]; self.J2_alpha[1][:] = self.J2_alpha[0]
        else:
            self.dLp0[:] = self.dLp2
            self.J2_beta[0][:] = self.J2_beta[2]; self.J2_alpha[0][:] = self.J2_alpha[2]

    def computeBondForceGeneral(self, plmode):
        L, N, nn = self.lib, self.N, self.nn
        if plmode == 6:
            self._geometry(self.dLp0)
            L.oracle_force(N, nn, 6, _p(self.neighbors), _p(self.nsign), _p(self.nb_initial), _p(self.Kn), _p(self.Tv), _p(self.damage_broken),
                           _p(self.dL), _p(self.dL_total), _p(self.TdL_total), _p(self.csx), _p(self.csy), _p(self.csz), _p(self.dL_ave),
                           _p(self.F), _p(self.Pin))
        elif plmode == 4:
            L.oracle_predictor(N, nn, _p(self.xyz), _p(self.xyz_temp), _p(self.neighbors), _p(self.nsign), _p(self.nb_initial), _p(self.Kn),
                               _p(self.Tv), _p(self.damage_broken), _p(self.F_temp), _p(self.csx), _p(self.csy), _p(self.csz), _p(self.ddL),
                               _p(self.ddL_total), _p(self.TddL_total), _p(self.F), _p(self.Pin))
        elif plmode == 0:
            self._geometry(self.dLp0)
            L.oracle_j2_return_map(N, nn, C.c_double(self.V), C.c_double(self.J2_H), C.c_double(self.J2_xi), _p(self.Ce), _p(self.type),
                                   _p(self.sigmay), _p(self.nsign), _p(self.nb), _p(self.nb_initial), _p(self.Kn), _p(self.Tv),
                                   _p(self.damage_w), _p(self.damage_broken), _p(self.distance_initial), _p(self.csx_initial),
                                   _p(self.csy_initial), _p(self.csz_initial), _p(self.dL), _p(self.dL_total), _p(self.TdL_total),
                                   _p(self.csx), _p(self.csy), _p(self.csz), _p(self.dLp0), _p(self.J2_beta[0]), _p(self.J2_alpha[0]),
                                   _p(self.dLp2), _p(self.J2_beta[2]), _p(self.J2_alpha[2]), _p(self.ddLp), _p(self.J2_dlambda),
                                   _p(self.pl_flag))
            self._geometry(self.dLp2)
            L.oracle_force(N, nn, 0, _p(self.neighbors), _p(self.nsign), _p(self.nb_initial), _p(self.Kn), _p(self.Tv), _p(self.damage_w),
                           _p(self.dL), _p(self.dL_total), _p(self.TdL_total), _p(self.csx), _p(self.csy), _p(self.csz), _p(self.dL_ave),
                           _p(self.F), _p(self.Pin))
        else:
            raise NotImplementedError(plmode)
        self.computeStress()
        self.switchStateV(2)

    def updateRR(self):
        rea = np.zeros(self.dim * self.N)
        k = self.lib.oracle_update_rr(self.N, self.dim, _p(self.dispBC_index), _p(self.Pex), _p(self.Pin), _p(self.residual), _p(rea))
        self.reaction_force = rea[:k]

    def calcStiffnessFiniteDifference(self):
        self.lib.oracle_fd_stiffness(self.N, self.nn, self.nconn, self.dim, C.c_double(self.radius), _p(self.xyz), _p(self.neighbors),
                                     _p(self.nsign), _p(self.nb_initial), _p(self.distance_initial), _p(self.dLp0), _p(self.damage_broken),
                                     _p(self.Kn), _p(self.Tv), _p(self.conn), _p(self.nb_conn), _p(self.kp0), _p(self.kp1), _p(self.K_global),
                                     _p(self.JK), _p(self.IK), _p(self.dL), _p(self.csx), _p(self.csy), _p(self.csz), _p(self.dL_total),
                                     _p(self.TdL_total), _p(self.F), _p(self.Pin))

    def setDispBC_stiffnessUpdate(self):
        self.lib.oracle_bc_stiffness_update(self.N, self.nconn, self.dim, _p(self.dispBC_index), _p(self.fix_index), _p(self.conn),
                                            _p(self.nb_conn), _p(self.kp0), _p(self.kp1), _p(self.K_global), _p(self.residual))

    def solverCG(self, rel=1e-8, abs_tol=1e-12):
        n = self.dim * self.N
        it = self.lib.oracle_cg(n, _p(self.IK), _p(self.JK), _p(self.K_global), _p(self.residual), _p(self.disp), C.c_double(rel),
                                C.c_double(abs_tol), n)
        self.xyz[:, :self.dim] += self.disp.reshape(self.N, self.dim)
        return it

    def updateDamageNonlocal(self, L, thr, Ac):
        return self.lib.oracle_damage_nonlocal(self.N, self.nn, C.c_double(L), C.c_double(thr), C.c_double(Ac), C.c_double(self.V),
                                               _p(self.xyz_initial), _p(self.neighbors), _p(self.nb_initial), _p(self.J2_dlambda),
                                               _p(self.J2_triaxiality), _p(self.damage_nonlocal0), _p(self.damage_broken), _p(self.damage_D0),
                                               _p(self.damage_w))

    def updateBrittleDamage(self, crit, nbreak, max_pairs=64):
        """constitutive.c:1437-1526; returns (candidate count, broken (i, neighbour) pairs in the reference's order)"""
        pairs = np.full((max_pairs, 2), -1, np.int32)
        self.lib.oracle_damage_brittle.restype = C.c_int
        k = self.lib.oracle_damage_brittle(self.N, self.nn, C.c_double(crit), int(nbreak), _p(self.neighbors), _p(self.nb_initial), _p(self.dL),
                                           _p(self.distance_initial), _p(self.damage_broken), _p(self.damage_D0), _p(self.damage_w), _p(pairs),
                                           max_pairs)
        return k, pairs[: min(k, int(nbreak)) if k > 0 else 0]

    def updateCrack(self):
        self.lib.oracle_update_crack(self.N, self.nn, self.dim, _p(self.nb_initial), _p(self.damage_broken), _p(self.damage_w), _p(self.csx),
                                     _p(self.csy), _p(self.csz), _p(self.F), _p(self.Pin), _p(self.nb), _p(self.damage_visual),
                                     _p(self.fix_index))
