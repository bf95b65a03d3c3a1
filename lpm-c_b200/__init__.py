"""lpm-c_b200 -- B200-native hot path of the nonlocal Lattice Particle Method (LPM-C).

The product is `liblpmb200.so` (hand-written sm_100a CUDA behind the C ABI in include/lpmb200.h)
plus the reference-named drop-in layer (csrc/dropin.c).  This Python package is a thin ctypes
host for tests and benchmarks; it contains no compute of its own and no CPU fallback.

The directory name carries a hyphen, so import it with
    import importlib; lpm = importlib.import_module("lpm-c_b200")
"""
from .capi import Context, LPMBError, lib, lib_path, device_count  # noqa: F401
from . import capi, lattice, driver  # noqa: F401
