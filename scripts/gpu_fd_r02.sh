#!/bin/bash
# FD assembly, round 2 (slice-cooperative kernels): parity tests, timing old vs new, ncu capture of the new kernels.
mkdir -p gpurun_out
if [ "$1" != "ncu" ]; then
timeout 600 python -m pytest tests/test_fd_variants_gpu.py tests/test_constitutive_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02v_fd_tests.log
timeout 300 python scripts/fd_profile.py 100 1,0 2>&1 | tee gpurun_out/r02v_fd_timing.log
timeout 600 python scripts/fd_profile.py 216 1,0 2>&1 | tee -a gpurun_out/r02v_fd_timing.log
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fd_rows -c 1 -f -o gpurun_out/r02v_fd_n100 python scripts/fd_profile.py 100 1 > gpurun_out/ncu_fd.log 2>&1
tail -3 gpurun_out/ncu_fd.log
