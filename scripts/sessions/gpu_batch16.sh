#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch16.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=8 run "loss step, CUDA graphs" 300 python scripts/loss_graph_time.py | tee gpurun_out/loss_graph_time.jsonl
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b16.json
