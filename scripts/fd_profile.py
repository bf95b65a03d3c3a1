"""FD tangent assembly on an n^3 SC block (default 100): timing, or a target for ncu (-k regex:fd_).
usage: fd_profile.py [n] [variants, e.g. 1,0]  (1 = slice-cooperative kernels, 0 = CTA-per-particle kernel)"""
import importlib
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402

lpm = importlib.import_module("lpm-c_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
c, info = bench.build_workload(lpm, n, 0, bricks=False)
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
for v in variants:
    c.set_params(fd_variant=float(v))
    for _ in range(3):
        t0 = time.time()
        c.fd_stiffness(False)
        c.synchronize()
        print(f"n={n}: fd_stiffness variant {v} {time.time() - t0:.4f} s", flush=True)
c.close()
