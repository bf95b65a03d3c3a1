"""pytest configuration.

-m "not gpu": oracle vs golden vectors / known answers, host logic, C-ABI symbol check (CPU only).
-m gpu      : parity tests proper -- they call the CUDA path through the C ABI (liblpmb200.so) and
              compare with the oracle / the committed golden vectors.  Nothing here reads
              /root/reference at run time; oracle/_ref/*.so (prebuilt from it) is used when present.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lpm():
    """the host package (ctypes over liblpmb200.so); importing fails loudly if the .so is missing"""
    return importlib.import_module("lpm-c_b200")


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "sc6_j2.npz")


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference compiled as a library (oracle/_ref); skip when it was not built"""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref/liblpmc_ref.so not built (needs /root/reference at build time)")
    r = oref.RefLPM.instance()
    r.threads(1)
    return r


@pytest.fixture
def ref_c1(ref):
    """default driver configuration C1 (21^3 SC) set up by the reference's own code, after the first
    FD assembly + BCs + predictor + updateRR of load step 1.  Function-scoped on purpose: the reference is one set of
    process globals that the tests advance (solves, load steps), so every user gets a fresh set-up (~3 s) and the
    known answers (80 then 106 CG iterations) do not depend on which tests ran before."""
    ref.setup_sc()
    nr, nf = ref.begin_step([(1, "z", 0.0)], [(2, 0.0, 0.0, -2000.0)])
    return {"ref": ref, "norm_residual": nr, "norm_reaction": nf}
