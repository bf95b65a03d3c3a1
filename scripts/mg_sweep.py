"""fast mode (param cg_precond): PCG iterations and time per solve for a few smoother settings on the bench workload.
python scripts/mg_sweep.py [n=216]"""
import importlib, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
lpm = importlib.import_module("lpm-c_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
c, info = bench.build_workload(lpm, n, 0, bricks=n >= 64)
c.set_dof_mask(c.get_field("dispBC_index"), c.get_field("fix_index"))
def solve(reps=3):
    ts, it = [], 0
    for _ in range(reps):
        c.copy_field("residual", "residual_save")
        c.synchronize()
        t0 = time.perf_counter()
        it, ok = c.solve_cg_device(update_xyz=False)
        c.synchronize()
        ts.append(time.perf_counter() - t0)
    return it, min(ts) * 1e3
it0, t0 = solve(2)
print(f"n={n}: plain CG {it0} iterations, {t0:.1f} ms per solve")
c.set_params(cg_precond=1.0)
for nu, om, om2 in ((2, 0.6, 0.6), (1, 0.7, 0.7), (2, 0.56, 1.39), (2, 0.5, 1.2), (2, 0.6, 1.0), (2, 0.55, 0.9), (2, 0.7, 0.5), (2, 0.8, 0.5), (2, 1.0, 0.5),
                    (3, 0.6, 1.0), (2, 0.6, 0.8), (4, 0.5, 1.2)):
    c.set_params(mg_nu=float(nu), mg_omega=om, mg_omega2=om2)
    try:
        it, ms = solve(3)
        print(f"  nu={nu} omega={om}/{om2}: {it} PCG iterations, {ms:.1f} ms per solve ({ms/max(it,1):.2f} ms per iteration)")
    except Exception as e:
        print(f"  nu={nu} omega={om}/{om2}: {str(e)[:80]}")
c.close()
