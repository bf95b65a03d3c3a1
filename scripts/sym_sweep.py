"""full-format vs symmetric-upper SpMV on real FD tangents (device timing); usage: python scripts/sym_sweep.py 100 160 216 [strip]"""
import importlib, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
lpm = importlib.import_module("lpm-c_b200")
sizes = [int(a) for a in sys.argv[1:] if int(a) > 32] or [100]
strips = [int(a) for a in sys.argv[1:] if int(a) <= 32] or [16]
for n in sizes:
    c, info = bench.build_workload(lpm, n, 0)
    N = n ** 3
    rng = np.random.default_rng(1)
    ms0 = c.spmv_bench(reps=10, variant=0)
    alg = c.spmv_bytes()
    out = {"n": n, "full_ms": round(ms0, 4), "full_GBs": round(alg / ms0 / 1e6, 1)}
    for strip in strips:
        c.set_param("sym_strip", strip)
        ms1 = c.spmv_bench(reps=10, variant=1)
        c.set_param("spmv_symmetric", 1)
        sb = c.spmv_bytes_stored()
        c.set_param("spmv_symmetric", 0)
        out[f"sym_ms_strip{strip}"] = round(ms1, 4)
        out[f"sym_speedup_strip{strip}"] = round(ms0 / ms1, 3)
        out["sym_stored_GB"] = round(sb / 1e9, 2)
        out[f"sym_GBs_of_stored_strip{strip}"] = round(sb / ms1 / 1e6, 1)
        break   # the schedule is built once per context
    if n <= 64:
        x = rng.standard_normal(3 * N)
        y0 = c.spmv(x)
        c.set_param("spmv_symmetric", 1)
        y1 = c.spmv(x)
        c.set_param("spmv_symmetric", 0)
        out["rel_diff"] = float(np.abs(y0 - y1).max() / np.abs(y0).max())
    print(json.dumps(out), flush=True)
    c.close()
