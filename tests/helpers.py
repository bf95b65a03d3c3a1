"""Shared helpers for the GPU parity tests: build a device context from the golden vectors and move
reference-layout state (as oracle/ref.py snapshots it) in and out of it."""
from __future__ import annotations

import numpy as np

BOND = ("dL", "dL_ave", "ddLp", "csx", "csy", "csz", "F", "bond_stress", "damage_broken", "damage_w")
PART = ("dL_total", "TdL_total", "stress_tensor", "J2_dlambda", "J2_stresseq", "J2_stressm", "J2_triaxiality", "xyz",
        "Pin", "pl_flag")


def params_from_golden(g) -> dict:
    return {str(k): float(v) for k, v in zip(g["param_names"], g["params"])}


def make_ctx(lpm, g, upload_initial_geometry=True):
    """context for the golden 6^3 case with topology, material and set-up state uploaded"""
    N, nn = g["setup.neighbors"].shape
    c = lpm.Context(N, 3, 2, nn, g["setup.conn"].shape[1])
    c.set_params(**params_from_golden(g))
    c.set_field("xyz", g["setup.xyz"])
    c.set_field("xyz_initial", g["setup.xyz"])
    if upload_initial_geometry:
        for n in ("distance_initial", "csx_initial", "csy_initial", "csz_initial"):
            c.set_field(n, g[f"setup.{n}"])
    c.set_neighbors(g["setup.neighbors"], g["setup.nsign"])
    c.set_connectivity(g["setup.conn"])
    c.set_field("type", g["setup.type"])
    c.set_field("sigmay", g["setup.sigmay"])
    c.calc_kntv(g["setup.Ce"])
    return c


def put_state(c, g, prefix, names=None):
    """upload a snapshot made by tests/golden/make_golden.py::state()"""
    for n in names or (BOND + PART):
        key = f"{prefix}.{n}"
        if key in g.files:
            c.set_field(n, g[key])
    if f"{prefix}.dLp" in g.files and names is None:
        put_slots(c, "dLp", g[f"{prefix}.dLp"])
        put_slots(c, "J2_beta", g[f"{prefix}.J2_beta"])
        put_slots(c, "damage_D", g[f"{prefix}.damage_D"])
        put_slots(c, "J2_alpha", g[f"{prefix}.J2_alpha"])
        put_slots(c, "damage_nonlocal", g[f"{prefix}.damage_nonlocal"])


def put_slots(c, name, arr):
    """reference [N][k][slots] (or [N][slots]) -> per-slot device fields name0, name1, ..."""
    arr = np.asarray(arr)
    for s in range(arr.shape[-1]):
        c.set_field(f"{name}{s}", np.ascontiguousarray(arr[..., s]))


def get_slots(c, name, nslots):
    return np.stack([c.get_field(f"{name}{s}") for s in range(nslots)], axis=-1)


def assert_same(a, b, what=""):
    """bit-exact (treating -0.0 == 0.0), with a useful message"""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        i = tuple(bad[0])
        denom = np.abs(b).max() or 1.0
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ; first at {i}: {a[i]!r} vs {b[i]!r}; "
                             f"max abs diff {np.abs(a - b).max():.3e} (scale {denom:.3e})")


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (n if n else 1.0))


def run_isolated(test_file, node, timeout=900):
    """Run ONE test body -- a function named impl_* of `test_file`, `node` = its name incl. any [param] -- in a child pytest
    process.  For GPU tests that have never been run: a crash of the library or a sticky CUDA error then stays in the child
    instead of taking the rest of the suite (or the pytest process itself) down."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-o", "python_functions=impl_*", "-m", "gpu",
                        f"{test_file}::{node}"], capture_output=True, text=True, timeout=timeout, cwd=root)
    assert r.returncode == 0, f"isolated run of {node} failed (rc {r.returncode}):\n{r.stdout[-3000:]}\n{r.stderr[-1500:]}"
    return r.stdout
