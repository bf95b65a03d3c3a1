"""CPU: the C-ABI library loads and exports every symbol include/lpmb200.h declares."""
import ctypes


def test_library_exports_every_declared_symbol(lpm):
    from importlib import import_module
    capi = import_module("lpm-c_b200.capi")
    names = capi.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(capi.lib, n)]
    assert not missing, f"declared in include/lpmb200.h but not exported: {missing}"


def test_no_cpu_fallback_without_device(lpm):
    """on a CPU-only host creating a context must fail loudly (no silent fallback)"""
    if lpm.device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    import pytest
    with pytest.raises(lpm.LPMBError) as e:
        lpm.Context(8, 3, 2, 18, 61)
    assert "no CUDA device" in str(e.value)


def test_version(lpm):
    assert lpm.lib.lpmb_version() >= 100


def test_dropin_exports_every_reference_entry_point():
    """liblpmc_dropin.so defines every function declared in include/lpmc_dropin.h (= the reference's stiffness.h,
    solver.h, constitutive.h).  It cannot be dlopen'ed on its own -- it references the driver's globals -- so the
    dynamic symbol table is read instead."""
    import re
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    so = root / "lpm-c_b200" / "liblpmc_dropin.so"
    assert so.exists(), "run __graft_entry__.build()"
    text = re.sub(r"/\*.*?\*/", "", (root / "include" / "lpmc_dropin.h").read_text(), flags=re.S)
    declared = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined"}
    out = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True, check=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = sorted(declared - defined)
    assert not missing, missing
    assert {"solverCG", "calcStiffness3DFiniteDifference", "computeBondForceGeneral", "updateRR", "switchStateV"} <= defined
    # and it needs the reference's globals from the driver it is linked into
    undef = subprocess.run(["nm", "-D", "--undefined-only", str(so)], capture_output=True, text=True, check=True).stdout
    assert " xyz" in undef and " K_global" in undef


def test_c_example_fails_loudly_without_device(lpm):
    """examples/sc_block (plain C over the C ABI) has no CPU path either"""
    import subprocess
    from pathlib import Path
    import pytest
    exe = Path(__file__).resolve().parents[1] / "examples" / "sc_block"
    if not exe.exists():
        pytest.skip("examples/sc_block not built")
    if lpm.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(exe), "8", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_every_entry_point_rejects_null_arguments_without_crashing():
    """error behaviour at the boundary: every function include/lpmb200.h declares, called with all-zero arguments (NULL
    context, NULL pointers), returns -- an error code for the calls that need a context, 0 for the pure getters -- instead of
    dereferencing NULL.  One child process walks all of them and prints each name before the call, so a crash names itself."""
    import re
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    hdr = re.sub(r"/\*.*?\*/", "", (root / "include" / "lpmb200.h").read_text(), flags=re.S)
    decls = re.findall(r"\b(lpmb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)
    assert len(decls) >= 50
    calls = [(n, 0 if a.strip() in ("", "void") else len(a.split(","))) for n, a in decls]
    code = (
        "import ctypes as C, sys\n"
        f"lib = C.CDLL({str(root / 'lpm-c_b200' / 'liblpmb200.so')!r})\n"
        f"for name, n in {calls!r}:\n"
        "    print('CALL', name, flush=True)\n"
        "    f = getattr(lib, name)\n"
        "    f.restype = C.c_longlong\n"
        "    rc = f(*([C.c_void_p(0)] * n))\n"
        "    print('RC', name, rc, flush=True)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("CALL")][-1]
    assert r.returncode == 0, f"crashed in {last}: {r.stderr[-500:]}"
    rcs = {ln.split()[1]: int(ln.split()[2]) for ln in r.stdout.splitlines() if ln.startswith("RC")}
    getters = {"lpmb_last_error", "lpmb_version", "lpmb_device_count", "lpmb_destroy", "lpmb_launch_count", "lpmb_stream",
               "lpmb_spmv_bytes_bricks", "lpmb_spmv_bytes", "lpmb_spmv_bytes_stored", "lpmb_dist_mode"}
    accepted_null = {n for n, rc in rcs.items() if rc == 0} - getters
    assert not accepted_null, f"accepted a NULL context: {sorted(accepted_null)}"


def test_c_multi_gpu_example_usage_and_loud_failure(lpm):
    """examples/sc_block_mgpu (one process per GPU from plain C): bad arguments -> usage, exit 2; without devices every rank
    says so and the parent reports the failure (no hang, no CPU path)"""
    import subprocess
    from pathlib import Path
    import pytest
    exe = Path(__file__).resolve().parents[1] / "examples" / "sc_block_mgpu"
    if not exe.exists():
        pytest.skip("examples/sc_block_mgpu not built")
    r = subprocess.run([str(exe), "3", "8"], capture_output=True, text=True, timeout=60)      # 8 layers cannot feed 3 ranks
    assert r.returncode == 2 and "usage" in r.stderr
    if lpm.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(exe), "2", "8", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and "rank 0 failed" in r.stderr


def test_python_binding_declares_argument_types_for_every_entry_point(lpm):
    """ctypes would otherwise pass 64-bit handles / sizes as C ints: every declared function that takes arguments has
    its argtypes set in lpm-c_b200/capi.py"""
    from importlib import import_module
    capi = import_module("lpm-c_b200.capi")
    no_args = {"lpmb_device_count", "lpmb_last_error", "lpmb_version"}
    missing = [n for n in capi.declared_symbols() if n not in no_args and getattr(capi.lib, n).argtypes is None]
    assert not missing, missing
