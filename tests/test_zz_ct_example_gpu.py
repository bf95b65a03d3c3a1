"""GPU integration: BASELINE.json config 5 AS SHIPPED -- examples/CT_sc_ductile_nonlocal.c (compact-tension specimen,
75 030 particles, 225 090 DoF, 18.4M upper non-zeros, J2 plasticity + nonlocal ductile damage, displacement control,
pre-cracked) -- linked UNCHANGED against the drop-in library (oracle/_ref/ct_sc_b200, built by oracle/Makefile from the
reference sources where they lie + the three global definitions the example forgot, oracle/shim/example_missing_globals.c).

Ground truth: tests/golden/c5src_*.txt = the first load steps of the all-CPU build of the same driver (oracle/_ref/ct_sc_cpu)
run SINGLE-THREADED (OMP_THREAD_LIMIT=1; 4-5 minutes per load step).  The reference's OpenMP loops race (SURVEY Appendix
D-1..3): with its own nt_force = 3 threads two runs of this very case needed 400 and 769 CG iterations for the first solve
and one of them diverged in step 3, so only the serial run is a usable oracle.  Nothing here reads /root/reference.
"""
import os
import re
import subprocess
import time
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
REFDIR = ROOT / "oracle" / "_ref"
GOLD = ROOT / "tests" / "golden"


def _table(path):
    rows = []
    for ln in Path(path).read_text().splitlines():
        try:
            rows.append([float(x) for x in ln.split()])
        except ValueError:
            continue
    return rows


def _log_counts(text):
    cg = [int(m) for m in re.findall(r"The system has been solved after (\d+) iterations", text)]
    newton = [int(m) for m in re.findall(r"Loading step \d+ has finished in (\d+) iterations", text)]
    return cg, newton


def _run_ct(tmp_path, nsteps, budget_s):
    """run oracle/_ref/ct_sc_b200 until the records of `nsteps` load steps are on disk (the driver itself would do 100)"""
    exe = REFDIR / "ct_sc_b200"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    t0 = time.time()
    with open(tmp_path / "run.log", "w") as log:
        p = subprocess.Popen(["stdbuf", "-oL", str(exe)], cwd=tmp_path, stdout=log, stderr=subprocess.STDOUT, env=env)
        try:
            while time.time() - t0 < budget_s and p.poll() is None:
                time.sleep(1.0)
                # "Time costed for step N" is printed after the step's records were written (driver :520)
                if f"Time costed for step {nsteps}:" in (tmp_path / "run.log").read_text():
                    break
        finally:
            if p.poll() is None:
                p.terminate()
                try:
                    p.wait(timeout=20)
                except subprocess.TimeoutExpired:
                    p.kill()
    text = (tmp_path / "run.log").read_text()
    assert "lpmc_dropin:" not in text, text[-2000:]
    assert "Particle number is 75030" in text
    cg, newton = _log_counts(text)
    assert len(newton) >= nsteps, f"only {len(newton)} load steps finished in {time.time() - t0:.0f} s\n" + text[-1500:]
    return cg, newton, time.time() - t0


def _compare_records(tmp_path, prefix, nsteps, tols):
    worst = {}
    for name, rtol in tols:
        a, b = _table(tmp_path / f"result_{name}.txt"), _table(GOLD / f"{prefix}_result_{name}.txt")
        n = min(len(a), len(b), nsteps + 1)
        assert n == nsteps + 1, (name, len(a), len(b))
        w = 0.0
        for k in range(n):
            ra, rb = np.array(a[k]), np.array(b[k])
            assert ra.shape == rb.shape, (name, k)
            # per-record scale: the largest component (stress records carry components that are zero by symmetry)
            w = max(w, float(np.abs(ra - rb).max() / max(np.abs(rb).max(), 1e-30)))
        worst[name] = w
        assert w <= rtol, f"{name}: worst relative record difference {w:.2e}"
    return worst


def test_ct_example_first_load_steps_match_serial_reference(tmp_path):
    """first 5 (elastic) load steps: Newton AND CG iteration counts identical, records to the printed digits"""
    gold_log = GOLD / "c5src_log.txt"
    if not (REFDIR / "ct_sc_b200").exists() or not gold_log.exists():
        pytest.skip("oracle/_ref/ct_sc_b200 or the golden records are missing")
    gcg, gnewton = _log_counts(gold_log.read_text())
    nsteps = len(gnewton)
    assert nsteps >= 2
    cg, newton, secs = _run_ct(tmp_path, nsteps, 420)
    # Newton iterations per load step: identical; CG iterations per solve: the stopping test ||r||^2 <= 1e-8 ||r0||^2 is
    # met within an iteration or two of the reference's (different summation order in the dot products)
    assert newton[:nsteps] == gnewton, (newton, gnewton)
    ncg = sum(gnewton)
    worst_cg = max(abs(a - b) for a, b in zip(cg[:ncg], gcg[:ncg]))
    assert worst_cg <= max(3, int(0.01 * max(gcg))), (cg[:ncg], gcg[:ncg])
    worst = _compare_records(tmp_path, "c5src", nsteps, (("force", 1e-6), ("disp", 1e-6), ("disp_CMOD1", 1e-6), ("disp_CMOD2", 1e-6), ("stress", 1e-5)))
    # broken-bond log: same TIMESTEP blocks (no bond breaks this early)
    bg = (tmp_path / "result_brokenbonds.txt").read_text().split("\n")
    bc = [ln for ln in (GOLD / "c5src_result_brokenbonds.txt").read_text().split("\n") if ln.strip()]
    assert bg[: len(bc)] == bc
    print(f"CT example: {nsteps} load steps, Newton {newton[:nsteps]}, CG {cg[:ncg]} vs {gcg[:ncg]}, record differences {worst}, {secs:.0f} s")


def test_ct_example_into_the_plastic_range(tmp_path):
    """34 load steps: the crack tip yields from step 28 on (30 particles through the thickness, Newton iterations 4 4 3 3
    4 4 5) and the nonlocal damage field grows to 0.12 -- the J2 return map, the three-slot state and the Gaussian damage
    gather of BASELINE config 5 on its own specimen.  Plastic steps that stop within a hair of the tolerance may take one
    Newton iteration more or less when rounding differs (the reference does that between thread counts); records to 1e-5.
    Golden: tests/golden/c5src_long_* (serial CPU run, ~55 minutes).  Written after the round's GPU budget was spent:
    the first B200 run of THIS test is the round-end suite (the 5-step test above ran green)."""
    gold_log = GOLD / "c5src_long_log.txt"
    if not (REFDIR / "ct_sc_b200").exists() or not gold_log.exists():
        pytest.skip("oracle/_ref/ct_sc_b200 or the golden records are missing")
    gcg, gnewton = _log_counts(gold_log.read_text())
    nsteps = len(gnewton)
    assert nsteps == 34 and max(gnewton) > 2
    cg, newton, secs = _run_ct(tmp_path, nsteps, 840)
    diff = [k for k in range(nsteps) if newton[k] != gnewton[k]]
    assert len(diff) <= 3 and all(abs(newton[k] - gnewton[k]) <= 1 for k in diff), (newton[:nsteps], gnewton)
    assert newton[:27] == gnewton[:27]                       # the elastic range is exact
    assert cg[: 2 * 27] == gcg[: 2 * 27] or max(abs(a - b) for a, b in zip(cg[:54], gcg[:54])) <= 4, (cg[:54], gcg[:54])
    worst = _compare_records(tmp_path, "c5src_long", nsteps, (("force", 1e-5), ("disp", 1e-5), ("disp_CMOD1", 1e-5), ("disp_CMOD2", 1e-5), ("stress", 1e-4)))
    bg = (tmp_path / "result_brokenbonds.txt").read_text().split("\n")
    bc = [ln for ln in (GOLD / "c5src_long_result_brokenbonds.txt").read_text().split("\n") if ln.strip()]
    assert bg[: len(bc)] == bc
    print(f"CT example: {nsteps} load steps, Newton {newton[:nsteps]} vs {gnewton}, record differences {worst}, {secs:.0f} s")
