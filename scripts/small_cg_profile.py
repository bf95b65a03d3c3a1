"""small-lattice regime (BASELINE configs 1-4 live here: matrix in L2, everything launch / latency bound): time the
device-resident CG and the whole Newton iteration on an n^3 block of the bench workload.  python scripts/small_cg_profile.py [n=21] [reps=20] [cg_graph=1]"""
import importlib, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
lpm = importlib.import_module("lpm-c_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c, info = bench.build_workload(lpm, n, 0, bricks=False)
if len(sys.argv) > 3:
    c.set_params(cg_graph=float(sys.argv[3]))    # 1 = batches of 16 iterations replayed as a CUDA graph (default), 0 = plain launches
for _ in range(3):
    bench.one_step(c)
c.synchronize()
t0 = time.perf_counter()
its = 0
for _ in range(reps):
    c.copy_field("xyz", "xyz_save")
    c.copy_field("residual", "residual_save")
    it, ok = c.solve_cg_device(update_xyz=False)
    its += it
c.synchronize()
t_cg = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for _ in range(reps):
    it2, nr = bench.one_step(c)
c.synchronize()
t_nw = (time.perf_counter() - t0) / reps
print(f"n={n} N={n**3}: CG solve {t_cg*1e3:.3f} ms for {its//reps} iterations = {t_cg*1e6/(its/reps):.2f} us per CG iteration; "
      f"whole Newton iteration {t_nw*1e3:.3f} ms (rest {1e3*(t_nw-t_cg):.3f} ms); fd_assembly {info['fd_assembly_s']*1e3:.1f} ms")
c.close()
