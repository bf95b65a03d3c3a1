"""SpMV / CG micro-sweep on synthetic SC blocks with the test-pattern matrix (device timing, CUDA events).
usage: python scripts/spmv_sweep.py 100 128 160"""
import importlib, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
lpm = importlib.import_module("lpm-c_b200")
peak = 6453.1
try:
    peak = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
for n in [int(a) for a in sys.argv[1:]] or [64, 100]:
    t0 = time.time()
    lat = lpm.lattice.sc_block(n)
    N = n ** 3
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_connectivity(lat["conn"])
    c.fill_test_pattern()
    t1 = time.time()
    ms = c.spmv_bench(reps=20)
    alg, sto = c.spmv_bytes(), c.spmv_bytes_stored()
    b = np.random.default_rng(1).standard_normal(3 * N)
    t2 = time.time()
    x, it, ok = c.solve_cg(b)
    t3 = time.time()
    print(json.dumps({"n": n, "N": N, "setup_s": round(t1 - t0, 2), "spmv_ms": round(ms, 4),
                      "alg_GBs": round(alg / ms / 1e6, 1), "stored_GBs": round(sto / ms / 1e6, 1),
                      "frac_of_measured_peak": round(alg / ms / 1e6 / peak, 3), "alg_MB": round(alg / 1e6, 1),
                      "cg_iters": it, "cg_wall_s": round(t3 - t2, 3), "launches": c.launches}), flush=True)
    c.close()
