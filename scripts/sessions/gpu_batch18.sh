#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch18.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=40 run "bi + topk" 600 python -m pytest tests/test_bi_gpu.py tests/test_topk_gpu.py -q
TAILN=12 run "suite" 900 python -m pytest tests -m gpu -q
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b18.json
