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


def test_c_multi_gpu_example_reproduces_default_case_known_answers(lpm):
    """examples/sc_block_mgpu: the default case C1 (21^3) on 2 GPUs from plain C over the C ABI (fork + exec per rank, NCCL id
    through a file): Newton iterations 2 2 1 and -- up to the summation order of the all-reduced dot products -- the CG
    iteration counts 80 / 106 of the single-GPU run"""
    import re
    if lpm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = ROOT / "examples" / "sc_block_mgpu"
    if not exe.exists():
        pytest.skip("examples/sc_block_mgpu not built")
    r = subprocess.run([str(exe), "2", "21", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    steps = re.findall(r"Loading step (\d+) has finished in (\d+) iterations; CG iterations:([ \d]+);", r.stdout)
    assert [int(s[1]) for s in steps] == [2, 2, 1], r.stdout
    cg1 = [int(x) for x in steps[0][2].split()]
    assert abs(cg1[0] - 80) <= 1 and abs(cg1[1] - 106) <= 1, cg1
    m = re.search(r"mean z-displacement of the loaded layer after 3 steps: (\S+)", r.stdout)
    assert m and abs(float(m.group(1)) / (3 * -1.27857453e-03) - 1.0) < 0.05      # ~linear in the elastic range


def test_two_gpu_slabs_with_lazy_halo_wait(lpm):
    """the experimental brick-by-brick halo wait (param brick_lazy_wait, LPMB_BRICK_LAZY_WAIT=1): same 2-GPU-vs-1-GPU check"""
    import os
    if lpm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, LPMB_BRICK_LAZY_WAIT="1")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dist_check_lite.py"), "40", "2"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 3])
def test_ct_specimen_on_slabs_matches_serial_reference(lpm, world):
    """BASELINE config 5 AS SHIPPED (examples/CT_sc_ductile_nonlocal.c: carved, pre-cracked compact-tension specimen, 75 030
    particles = 15 z-layers of 5 002) on 2 and 3 z-slabs over the C ABI (tests/dist_ct_check.py): Newton iteration counts, the
    CG iteration count of every solve and the printed residual / reaction norms of three load steps equal the SERIAL all-CPU
    run of the unchanged example (tests/golden/c5src_log.txt); the device-built neighbour lists of every slab equal the
    reference's.  Geometry / types / initial crack come from the reference's host code (oracle/_ref) in the parent."""
    from oracle import ref as oref
    if lpm.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if not oref.available():
        pytest.skip("oracle/_ref/liblpmc_ref.so not built")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dist_ct_check.py"), str(world), "3"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CT_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    print(r.stdout[-600:])
