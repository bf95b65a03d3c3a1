#!/bin/bash
# round 2, second session: sanitizer on the new FD kernels, ncu capture of the fused J2 passes
mkdir -p gpurun_out; out=gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_fd_variants_gpu.py > $out/r02x_memcheck_fd.log 2>&1; echo "memcheck fd rc=$?" | tee -a $out/r02x_memcheck_fd.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_fd_variants_gpu.py -k "chunked or repeatable" > $out/r02x_racecheck_fd.log 2>&1; echo "racecheck fd rc=$?" | tee -a $out/r02x_racecheck_fd.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_fd_variants_gpu.py -k "chunked or repeatable" > $out/r02x_synccheck_fd.log 2>&1; echo "synccheck fd rc=$?" | tee -a $out/r02x_synccheck_fd.log
timeout 300 python scripts/j2_profile.py 216 2>&1 | tee $out/r02x_j2_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:j2_ -c 2 -f -o $out/r02x_j2_n100 python scripts/j2_profile.py 100 > $out/ncu_j2.log 2>&1
tail -2 $out/ncu_j2.log
