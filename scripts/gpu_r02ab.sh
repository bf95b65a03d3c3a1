#!/bin/bash
# J2 fused kernel with shared-memory per-bond caches: parity tests, timing, ncu; racecheck of the warp-per-particle cp kernel
mkdir -p gpurun_out; out=gpurun_out
timeout 600 python -m pytest tests/test_constitutive_gpu.py tests/test_trajectory_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee $out/r02ab_j2_tests.log
timeout 300 python scripts/j2_profile.py 216 2>&1 | tee $out/r02ab_j2_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:j2_ -c 2 -f -o $out/r02ab_j2_n100 python scripts/j2_profile.py 100 > $out/ncu_j2.log 2>&1
tail -2 $out/ncu_j2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_cp_gpu.py -k "warp_kernel or fcc_topology" > $out/r02ab_racecheck_cp.log 2>&1; echo "racecheck cp rc=$?" | tee -a $out/r02ab_racecheck_cp.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_cp_gpu.py -k "warp_kernel or two_load" > $out/r02ab_memcheck_cp.log 2>&1; echo "memcheck cp rc=$?" | tee -a $out/r02ab_memcheck_cp.log
