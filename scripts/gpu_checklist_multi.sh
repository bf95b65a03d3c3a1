#!/usr/bin/env bash
# Multi-GPU re-validation + scaling line after the tile trimming (never measured at 4 / 8 GPUs in round 1).
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_checklist_multi.sh r02m 8'
# (charged N x box time; at N = 2 it takes ~3 minutes).  Outputs: gpurun_out/<tag>_*.log
set -u
tag=${1:-r02m}
maxn=${2:-2}
out=gpurun_out
mkdir -p "$out"
py=python
port=29530
# correctness first: slabs == single GPU, with the library's own rendezvous (no torch) and through torch.distributed
for w in 2 4; do
  [ "$w" -le "$maxn" ] || continue
  timeout 600 $py tests/dist_check_lite.py 40 "$w" > "$out/${tag}_dist_lite_w${w}.log" 2>&1
  echo "rc=$?" >> "$out/${tag}_dist_lite_w${w}.log"
done
timeout 900 $py -m pytest tests/test_dist_gpu.py -m gpu -q -rxX > "$out/${tag}_dist_tests.log" 2>&1
# plain-C multi-GPU driver (never run in round 1): default case on 2 ranks, then the 10M-particle case on all of them
timeout 300 examples/sc_block_mgpu 2 21 3 > "$out/${tag}_sc_block_mgpu_21.log" 2>&1
echo "rc=$?" >> "$out/${tag}_sc_block_mgpu_21.log"
timeout 600 examples/sc_block_mgpu "$maxn" 216 1 > "$out/${tag}_sc_block_mgpu_216.log" 2>&1
echo "rc=$?" >> "$out/${tag}_sc_block_mgpu_216.log"
# scaling: the bench exactly as the driver launches it
timeout 600 $py bench.py --no-cpu-baseline > "$out/${tag}_bench_n1.log" 2>&1
for w in 2 4 8; do
  [ "$w" -le "$maxn" ] || continue
  port=$((port + 1))
  timeout 900 $py -m torch.distributed.run --nnodes=1 --nproc-per-node "$w" --master-addr 127.0.0.1 --master-port "$port" \
    bench.py --gpus "$w" --steps 3 --warmup 3 > "$out/${tag}_bench_n${w}.log" 2>&1
  # the same without the tile trimming, for the A/B
  port=$((port + 1))
  timeout 900 $py -m torch.distributed.run --nnodes=1 --nproc-per-node "$w" --master-addr 127.0.0.1 --master-port "$port" \
    bench.py --gpus "$w" --steps 3 --warmup 3 --no-brick-trim > "$out/${tag}_bench_notrim_n${w}.log" 2>&1
done
# experimental: halo wait brick by brick (interior layers first); correctness, then the scaling lines
w=$maxn
LPMB_BRICK_LAZY_WAIT=1 timeout 600 $py tests/dist_check_lite.py 40 2 > "$out/${tag}_dist_lite_lazy_w2.log" 2>&1
echo "rc=$?" >> "$out/${tag}_dist_lite_lazy_w2.log"
port=$((port + 1))
LPMB_BRICK_LAZY_WAIT=1 timeout 900 $py -m torch.distributed.run --nnodes=1 --nproc-per-node "$w" --master-addr 127.0.0.1 --master-port "$port" \
  bench.py --gpus "$w" --steps 3 --warmup 3 > "$out/${tag}_bench_lazy_n${w}.log" 2>&1
grep -h '"metric"' "$out"/${tag}_bench_n*.log "$out"/${tag}_bench_lazy_n*.log "$out"/${tag}_bench_notrim_n*.log | $py -c '
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    print(d["n_gpus"], "GPUs:", round(d["value"], 4), d["unit"], "| slowest-rank SpMV", round(d["roofline"]["avg_launch_ms"], 4), "ms =",
          round(d["roofline"]["frac"], 3), "of peak | e2e", round(d["e2e"]["value"], 4))
' | tee "$out/${tag}_scaling_summary.txt"
