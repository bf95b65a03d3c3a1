"""diagnostic: where do F / Pin of the drop-in run differ from the reference run on the bench sample?"""
import sys, os, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
import argparse
m = int(sys.argv[1]) if len(sys.argv) > 1 else 48
args = argparse.Namespace(cpu_sample_n=m, cpu_steps=2, cpu_threads=0, dropin_s1=0)
with bench._QuietStdout():
    smp = bench.reference_newton_sample(m, 2, 1, bench.usable_cores(), keep_outputs=True)
cpu = smp.pop("arrays")
print("cpu", {k: v for k, v in smp.items()})
g, garr = bench.run_dropin_sample(args, m, 2, 0)
print("gpu", g)
N = m ** 3
for k in ("disp", "F", "Pin", "xyz"):
    a, b = garr[k], cpu[k]
    d = np.abs(a - b).reshape(N, -1).max(axis=1)
    bad = np.nonzero(d > 1e-9 * np.abs(b).max())[0]
    print(k, "rel", bench._rel(a, b), "max abs", d.max(), "scale", np.abs(b).max(), "bad particles", bad.size)
    if bad.size:
        z = bad // (m * m)
        print("   bad z-layers histogram:", np.bincount(z, minlength=m).tolist())
        print("   first bad:", bad[:10].tolist(), "values gpu", a.reshape(N, -1)[bad[0]][:6], "cpu", b.reshape(N, -1)[bad[0]][:6])
# the reference once more with the bond-force loop on ONE thread
from oracle.ref import RefLPM
r = RefLPM.instance()
r.put("xyz", cpu["xyz0"])
L = r.lib
def bf():
    L.omp_set_num_threads(1)
    L.computeBondForceGeneral(0, 1)
    L.omp_set_num_threads(bench.usable_cores())
with bench._QuietStdout():
    # residual was overwritten by updateRR; the solve is not repeated: only the law on the final positions
    r.put("xyz", cpu["xyz"])
    L.switchStateV(0)
    bf()
F1, P1 = r.get("F"), r.get("Pin")
print("reference 1-thread law vs reference threaded law: F", bench._rel(cpu["F"], F1), "Pin", bench._rel(cpu["Pin"], P1))
print("gpu vs reference 1-thread law: F", bench._rel(garr["F"], F1), "Pin", bench._rel(garr["Pin"], P1))
