"""CPU: bench.py's reference arm (the reference's own CPU code from oracle/_ref, timed on the host cores) prints exactly one
JSON line with the keys the driver reads; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-n", "12",
                        "--cpu-steps", "1", "--ref-s1-n", "14", "--ref-s1-steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                       # the reference's own printf output stays off stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "newton_iterations_per_second" and d["unit"] == "Newton it/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ms_per_step is the MEASURED time of one timed sample step; the S1 anchor and the same-config sample are reported beside it
    assert d["ms_per_step"] > 0 and d["ms_per_step_config"] == pytest.approx(1000.0 / d["value"])
    assert d["sample_measured"]["particles"] == 12 ** 3 and d["s1_measured"]["particles"] == 14 ** 3
    assert d["same_config_sample"]["particles"] == 12 ** 3 and d["same_config_sample"]["cpu_it_per_s"] > 0


def test_gpu_arm_fails_loudly_without_device(lpm):
    if lpm.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--lattice-n", "8"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]       # no number without the CUDA path
