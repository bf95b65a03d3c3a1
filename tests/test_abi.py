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
