#!/usr/bin/env bash
# one 8-GPU call (charged 8x): scaling line with dist_parity, lazy-halo-wait A/B, C5 load steps end to end on 8 slabs
set -u
out=gpurun_out; mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); run 29521 --steps 5 --warmup 3" > $out/r02h_bench_n8.log 2> $out/r02h_bench_n8.err
tail -c 2500 $out/r02h_bench_n8.log; tail -2 $out/r02h_bench_n8.err
LPMB_BRICK_LAZY_WAIT=1 timeout 300 bash -c "$(declare -f run); run 29522 --steps 5 --warmup 3 --dist-parity-n 0" > $out/r02h_bench_n8_lazy.log 2> $out/r02h_bench_n8_lazy.err
python - <<PY
import json
for f in ("r02h_bench_n8.log","r02h_bench_n8_lazy.log"):
    try:
        d=json.loads([l for l in open("gpurun_out/"+f) if l.startswith("{")][-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "spmv", [round(r["spmv_ms"],4) for r in d["roofline"]["per_rank"]], "parity", (d.get("dist_parity") or {}).get("ok"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 400 ./examples/sc_block_mgpu 8 216 4 c5 0.005 2>&1 | tail -16 | tee $out/r02h_sc_block_mgpu_c5_n216_8gpu.log
