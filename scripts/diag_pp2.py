"""diagnostic: regenerate sc6_particle2.npz through the drop-in layer and print every relative error vs the golden file"""
import os, subprocess, sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
out = Path(tempfile.mkdtemp()) / "pp2.npz"
env = dict(os.environ, LPMB_REF_SO=str(ROOT / "oracle/_ref/liblpmc_b200host.so"), LPMB_GOLDEN_OUT=str(out))
r = subprocess.run([sys.executable, str(GOLD / "make_golden_particle2.py")], env=env, capture_output=True, text=True, timeout=600)
print("rc", r.returncode); print(r.stdout[-3000:]); print(r.stderr[-3000:])
if r.returncode: sys.exit(1)
new, old = np.load(out), np.load(GOLD / "sc6_particle2.npz")
def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
for k in sorted(old.files):
    if k not in new.files: print("MISSING", k); continue
    if old[k].dtype.kind in "fi" and old[k].shape == new[k].shape:
        e = rel(new[k], old[k])
        if e > 1e-12: print(f"{k:40s} rel {e:.3e} maxabs {np.abs(np.asarray(new[k],float)-np.asarray(old[k],float)).max():.3e} nbad {(new[k]!=old[k]).sum()}")
    else: print("shape/dtype", k, old[k].shape, new[k].shape)
