#!/usr/bin/env bash
# round-2 evidence on one B200: fast-mode tests + smoother sweep, ncu launch list + full captures of the dominant kernels
# (never bench values), compute-sanitizer on the kernels added / changed this round.  Outputs: gpurun_out/r02i_*
set -u
out=gpurun_out; mkdir -p $out; tag=r02i
timeout 300 python -m pytest tests/test_solver_gpu.py -m gpu -q -s -k "fast_mode" > $out/${tag}_fast_tests.log 2>&1; tail -6 $out/${tag}_fast_tests.log
timeout 400 python scripts/mg_sweep.py 216 > $out/${tag}_mg_sweep_n216.log 2>&1; cat $out/${tag}_mg_sweep_n216.log
B="python bench.py --steps 1 --warmup 2 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file $out/${tag}_launches_n216.csv $B > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"brick_spmv|brick_gather" -s 20 -c 2 -o $out/${tag}_brick_n216 -f $B --no-fast-mode > $out/${tag}_ncu_brick.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mg_stencil_tiled" -s 8 -c 2 -o $out/${tag}_mg_n216 -f $B > $out/${tag}_ncu_mg.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fd_stiffness|symmetrize|j2_return_map|geometry_kernel|force_kernel|stress_kernel|update_rr|nonlocal_damage" -c 12 -o $out/${tag}_constitutive_n216 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast-mode > $out/${tag}_ncu_const.log 2>&1
for r in brick mg constitutive; do ncu -i $out/${tag}_${r}_n216.ncu-rep --page raw --csv > $out/${tag}_${r}_n216_ncu_raw.csv 2>/dev/null; done
ls -la $out | grep $tag
if command -v compute-sanitizer > /dev/null; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_solver_gpu.py -k "brick or fast_mode or cg" > $out/${tag}_memcheck_solver.log 2>&1; echo "memcheck solver rc=$?" | tee -a $out/${tag}_memcheck_solver.log
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_variants_gpu.py tests/test_zzz_particle2_gpu.py -k "particle_laws or per_particle_j2_energy_and_iso_laws or damage or per_particle_crystal" > $out/${tag}_memcheck_variants.log 2>&1; echo "memcheck variants rc=$?" | tee -a $out/${tag}_memcheck_variants.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_solver_gpu.py -k "stream_only_needed_rows or fast_mode" > $out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $out/${tag}_racecheck.log
  timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_solver_gpu.py -k "fast_mode" > $out/${tag}_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $out/${tag}_synccheck.log
fi
