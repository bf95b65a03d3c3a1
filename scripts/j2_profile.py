"""One J2 constitutive call (computeBondForceGeneral(0): j2_fused_kernel + j2_force_stress_kernel) on an n^3 SC block in
the plastic range: timing with CUDA-synchronised wall clock, or a target for ncu (-k regex:j2_).
usage: j2_profile.py [n]"""
import importlib
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402

lpm = importlib.import_module("lpm-c_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
c, info = bench.build_workload(lpm, n, 0, bricks=False)
bench.one_step(c)          # one Newton iteration: the state is plastic afterwards
c.synchronize()
for _ in range(3):
    t0 = time.time()
    c.bond_force(0)
    c.synchronize()
    print(f"n={n}: computeBondForceGeneral(0) {1e3 * (time.time() - t0):.3f} ms", flush=True)
c.close()
