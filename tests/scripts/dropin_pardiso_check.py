"""Child process of tests/test_dropin_gpu.py::test_dropin_solver_pardiso: the reference's HOST code (oracle/_ref/
liblpmc_b200host.so = its driver TU + boundary.c / initialization.c / neighbor.c ... linked against liblpmc_dropin.so)
sets up the golden 6^3 case, assembles the tangent, applies the displacement BCs to K_global / residual on the host
(boundary.c:159-281) and calls solverPARDISO() (solver.h:5, solver.c:3-92) through the drop-in layer.  Check: the
displacement solves the BC-modified system -- dense LU of the reference's own K_global / IK / JK in numpy -- to 1e-9,
and xyz moved by exactly disp (solver.c:88-91)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.environ["LPMB_REF_SO"] = str(ROOT / "oracle" / "_ref" / "liblpmc_b200host.so")
from oracle.ref import RefLPM  # noqa: E402

r = RefLPM.instance()
r.threads(1)
r.setup_sc(box=(-0.2, 2.7, -0.2, 2.7, -0.2, 2.7), top_z=2.5, plmode=0)
L = r.lib
N, dim = r.N, r.dim
assert N == 216
r.begin_step([(2, "x", 0.0), (2, "y", 0.0), (2, "z", 0.0)], [(1, 0.0, 0.0, -2000.0)])   # bottom layer clamped: K is non-singular
L.switchStateV(0)
L.setDispBC_stiffnessUpdate3D()
K, IK, JK = r.get("K_global"), r.get("IK"), r.get("JK")
rhs = r.get("residual")
n = dim * N
A = np.zeros((n, n))
for row in range(n):
    for e in range(IK[row] - 1, IK[row + 1] - 1):
        A[row, JK[e] - 1] = K[e]
A = A + np.triu(A, 1).T
want = np.linalg.solve(A, rhs)
x0 = r.get("xyz")
L.solverPARDISO()
disp = r.get("disp")
err = float(np.linalg.norm(disp - want) / np.linalg.norm(want))
moved = float(np.abs((r.get("xyz") - x0).ravel() - disp).max())
print(f"PARDISO_CHECK rel_err {err:.3e} xyz_minus_disp {moved:.3e}")
assert err <= 1e-9, err
assert moved <= 1e-15 * max(1.0, float(np.abs(x0).max())), moved
# the loose reference CG (1e-8 on squared norms) would be ~1e-5 off: the direct-solve stand-in must be much tighter
print("PARDISO_CHECK OK")
