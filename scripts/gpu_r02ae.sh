#!/bin/bash
# CUDA-graph batches in the small-lattice CG: tests, per-iteration time, the default driver's wall time
mkdir -p gpurun_out; out=gpurun_out
timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_trajectory_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee $out/r02ae_solver_tests.log
for g in 1 0; do for n in 21 41; do timeout 200 python scripts/small_cg_profile.py $n 20 $g 2>&1 | tail -1 | sed "s/^/cg_graph=$g: /"; done; done | tee $out/r02ae_small_cg.log
cd /tmp && rm -rf c1run && mkdir c1run && cd c1run && ( time LPMB_DROPIN_PROFILE=1 $GRAFT_REPO_ROOT/oracle/_ref/lpmc_default_b200 > run.log 2>&1 ) 2>&1 | tail -4 | tee $GRAFT_REPO_ROOT/$out/r02ae_c1_dropin_profile.log; grep -A16 "lpmc_dropin profile" run.log | tee -a $GRAFT_REPO_ROOT/$out/r02ae_c1_dropin_profile.log
