"""crystal-plasticity law (plmode 1, cp_miehe_kernel) timing on FCC blocks of BASELINE config 4's material:
   python scripts/cp_profile.py [cells=12 -> 6 912 particles] [reps=3] [cp_warp=1]        (63 cells -> 1 000 188 particles)"""
import importlib, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
lpm = importlib.import_module("lpm-c_b200")
g = np.load(ROOT / "tests" / "golden" / "fcc_cp.npz")
par = {str(k): float(v) for k, v in zip(g["param_names"], g["params"])}
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
r = par["radius"]
a = 2.0 * np.sqrt(2.0) * r                      # FCC cell edge for nearest-neighbour distance 2 r
base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
ijk = np.stack(np.meshgrid(np.arange(cells), np.arange(cells), np.arange(cells), indexing="ij"), -1).reshape(-1, 3)
xyz = (ijk[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
order = np.lexsort((xyz[:, 0], xyz[:, 1], xyz[:, 2]))      # x fastest, z slowest like the reference
xyz = np.ascontiguousarray(xyz[order])
N = xyz.shape[0]
S = int(par["nslipSys"])
c = lpm.Context(N, 3, 3, 18, 61)
c.set_params(**{k: v for k, v in par.items() if k != "nslipSys"})
c.set_field("xyz", xyz)
c.set_field("xyz_initial", xyz)
t0 = time.perf_counter()
c.build_topology(par["neighbor1_cutoff"], par["neighbor2_cutoff"])
c.set_field("type", np.zeros(N, dtype=np.int32))
c.calc_kntv(g["setup.Ce"][:1])
c.compute_dl()
c.set_schmid_tensor(g["setup.schmid_tensor"])
gy0 = np.broadcast_to(g["setup.cp_gy"][0], (N, S, 3))
for s in range(3):
    c.set_field(f"cp_gy{s}", np.ascontiguousarray(gy0[..., s]))
c.synchronize(); t1 = time.perf_counter()
c.compute_cab()
c.synchronize(); t2 = time.perf_counter()
nb = c.get_field("nb_initial") if False else None
x1 = xyz.copy(); x1[:, 2] *= 1.0 + 4e-4; x1[:, 0] *= 1.0 - 1.2e-4; x1[:, 1] *= 1.0 - 1.2e-4   # uniaxial stretch beyond yield
c.set_field("xyz", x1)
if len(sys.argv) > 3:
    c.set_params(cp_warp=float(sys.argv[3]))     # 1 = warp-per-particle kernel (default), 0 = thread-per-particle kernel
ts = []
for _ in range(reps):
    c.switch_state(0)
    c.synchronize(); a0 = time.perf_counter()
    c.bond_force(1)
    c.synchronize(); ts.append(time.perf_counter() - a0)
act = int(c.get_field("cp_Jact").sum())
print(f"FCC {cells}^3 cells = {N} particles, {S} slip systems: topology+set-up {t1-t0:.2f} s, computeCab {t2-t1:.3f} s, "
      f"computeBondForceGeneral(1) {min(ts)*1e3:.2f} ms (best of {reps}; {min(ts)/N*1e9:.1f} ns per particle), active systems {act} "
      f"({act/N:.2f} per particle)")
c.close()
