"""one fast-mode solve (multigrid PCG) between cudaProfilerStart/Stop: run under
   ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ... python scripts/fast_profile.py [n=216]"""
import ctypes, importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
lpm = importlib.import_module("lpm-c_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
c, info = bench.build_workload(lpm, n, 0, bricks=n >= 64)
c.set_dof_mask(c.get_field("dispBC_index"), c.get_field("fix_index"))
c.set_params(cg_precond=1.0)
rt = None
for name in ("libcudart.so", "libcudart.so.12"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
def solve():
    c.copy_field("residual", "residual_save")
    return c.solve_cg_device(update_xyz=False)
whole_step = len(sys.argv) > 2 and sys.argv[2] == "step"   # profile a whole Newton iteration (bench.one_step) instead of the solve alone
solve()
bench.one_step(c)
c.synchronize()
if rt: rt.cudaProfilerStart()
if whole_step:
    it, nr = bench.one_step(c)
    ok = True
else:
    it, ok = solve()
c.synchronize()
if rt: rt.cudaProfilerStop()
print("fast-mode", "Newton iteration:" if whole_step else "solve:", it, "PCG iterations", ok)
if len(sys.argv) > 3 and sys.argv[3] == "breakdown":
    import time
    def timed(label, fn, reps=3):
        c.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        c.synchronize(); print(f"  {label:42s} {(time.perf_counter()-t0)/reps*1e3:9.3f} ms")
    for mode in (1.0, 0.0):
        c.set_params(cg_precond=mode)
        print("cg_precond =", mode)
        timed("one_step (restore + newton_iteration)", lambda: bench.one_step(c))
        timed("copy_field xyz + residual", lambda: (c.copy_field("xyz", "xyz_save"), c.copy_field("residual", "residual_save")))
        timed("switch_state(0)", lambda: c.switch_state(0))
        timed("solve_cg_device(update_xyz=True)", lambda: (c.copy_field("residual", "residual_save"), c.solve_cg_device(update_xyz=True)), reps=2)
        timed("bond_force(0)", lambda: c.bond_force(0))
        timed("update_rr", lambda: c.update_rr())
c.close()
