#!/usr/bin/env bash
# One gpurun call that re-validates everything written without a GPU at the end of round 1 and refreshes the evidence
# under profiles/.  Usage (from the repo root, ~12 GPU-minutes on one B200):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_checklist.sh r02a'
# Outputs land in gpurun_out/<tag>_*.log (copy what is to be judged into profiles/).
set -u
tag=${1:-r02a}
out=gpurun_out
mkdir -p "$out"
py=python

# 1. the whole GPU suite in the driver's order (includes the hedged tests: BCC crystal plasticity, the 34-step CT run,
#    the plmode 3 / 5 per-particle entry points in tests/test_zzz_particle2_gpu.py)
timeout 1200 $py -m pytest tests -m gpu -x -q -rxX --durations=15 > "$out/${tag}_gpu_suite.log" 2>&1
echo "gpu suite rc=$?" | tee -a "$out/${tag}_gpu_suite.log"

# 2. default bench line (N = 1) and the full-format kernel beside it
timeout 600 $py bench.py > "$out/${tag}_bench.log" 2>&1
timeout 300 $py bench.py --spmv full --no-cpu-baseline --steps 2 > "$out/${tag}_bench_full.log" 2>&1

# 3. sanitizer on the kernels added after the last sanitizer run (small cases only)
if command -v compute-sanitizer > /dev/null; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 $py -m pytest -q -x -m gpu \
    tests/test_zzz_particle2_gpu.py tests/test_variants_gpu.py -k "particle or damage or 2d or bcc" > "$out/${tag}_memcheck.log" 2>&1
  echo "memcheck rc=$?" >> "$out/${tag}_memcheck.log"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $py -m pytest -q -x -m gpu tests/test_solver_gpu.py -k "brick" \
    > "$out/${tag}_memcheck_brick.log" 2>&1
  echo "memcheck rc=$?" >> "$out/${tag}_memcheck_brick.log"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 $py -m pytest -q -x -m gpu tests/test_solver_gpu.py \
    -k "stream_only_needed_rows" > "$out/${tag}_racecheck_brick.log" 2>&1
  echo "racecheck rc=$?" >> "$out/${tag}_racecheck_brick.log"
fi

# 4. launch list + one full capture of the brick SpMV on the final build (never a bench value)
if command -v ncu > /dev/null; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file "$out/${tag}_launches_n216.csv" \
    $py bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_ncu_launches.log" 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_ -s 40 -c 2 -o "$out/${tag}_brick_n216" -f \
    $py bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_ncu_full.log" 2>&1
fi
ls -la "$out" | tail -20
