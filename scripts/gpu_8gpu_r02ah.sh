#!/usr/bin/env bash
# one 8-GPU call (charged 8x): scaling line (parity mode + dist_parity + fast mode on slabs), C5 load steps end to end in fast mode
set -u
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); run 29521 --steps 3 --warmup 3" > $out/r02ah_bench_n8.log 2> $out/r02ah_bench_n8.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02ah_bench_n8.log") if l.startswith("{")][-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", (d.get("dist_parity") or {}).get("ok"), "fast", d.get("fast_mode"))
except Exception as e: print("ERR", e)
PY
tail -2 $out/r02ah_bench_n8.err
timeout 200 ./examples/sc_block_mgpu 8 216 4 c5 0.005 fast 2>&1 | tail -14 | tee $out/r02ah_sc_block_mgpu_c5_n216_8gpu_fast.log
