#!/bin/bash
# round 2, second session: warp-per-particle crystal-plasticity kernel (tests + timing), brittle examples
mkdir -p gpurun_out; out=gpurun_out
timeout 600 python -m pytest tests/test_cp_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee $out/r02z_cp_tests.log
for w in 1 0; do timeout 200 python scripts/cp_profile.py 12 3 $w 2>&1 | tail -1; done | tee $out/r02z_cp_timing.log
for w in 1 0; do timeout 300 python scripts/cp_profile.py 63 3 $w 2>&1 | tail -1; done | tee -a $out/r02z_cp_timing.log
timeout 600 python -m pytest tests/test_dropin_gpu.py -x -q -m gpu -s -k "brittle or config4" 2>&1 | tail -12 | tee $out/r02z_brittle_tests.log
