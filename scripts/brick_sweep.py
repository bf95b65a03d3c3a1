"""SpMV: full-format SELL kernel vs brick-blocked symmetric kernel (lpmb_brick.cu) on simple-cubic blocks.
usage: python scripts/brick_sweep.py [n ...]   (run on the GPU box)"""
import importlib
import sys
import time

import numpy as np

sys.path.insert(0, ".")
lpm = importlib.import_module("lpm-c_b200")

for n in [int(a) for a in sys.argv[1:]] or [48, 104, 216]:
    lat = lpm.lattice.sc_block(n)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25)
    c.set_field("xyz_initial", lat["xyz"])
    c.set_connectivity(lat["conn"])
    del lat
    c.fill_test_pattern()
    t_full = c.spmv_bench(20, 0)
    t0 = time.time()
    c.enable_bricks(True)
    t_b = c.spmv_bench(20, 2)
    setup = time.time() - t0
    bf, bb = c.spmv_bytes(), c.spmv_bytes_bricks()
    line = f"n={n} N={N} full {t_full:.3f} ms ({bf / t_full / 1e6:.0f} GB/s)  brick {t_b:.3f} ms ({bb / t_b / 1e6:.0f} GB/s of {bb / 1e9:.2f} GB)  speedup {t_full / t_b:.3f}  setup+first {setup:.2f} s"
    if n <= 104:
        rng = np.random.default_rng(1)
        x = rng.standard_normal(3 * N)
        y1 = c.spmv(x)
        c.enable_bricks(False)
        y0 = c.spmv(x)
        line += f"  rel diff {np.abs(y0 - y1).max() / np.abs(y0).max():.2e}"
    print(line, flush=True)
    c.close()
