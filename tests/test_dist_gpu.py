"""GPU (>= 2 devices): slab-decomposed Newton iteration == single-GPU Newton iteration."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_gpu_slabs_match_single_gpu(lpm):
    if lpm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(ROOT / "tests" / "dist_check.py"), "24"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_gpu_slabs_match_single_gpu_without_torch(lpm):
    """the same check with the library's own rendezvous only (NCCL id through a file, system libnccl): no torch in any
    process -- what a C host would do (tests/dist_check_lite.py)"""
    if lpm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dist_check_lite.py"), "24", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
