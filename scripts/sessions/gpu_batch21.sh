#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_batch21.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=30 run "topk + bi + sharded" 600 python -m pytest tests/test_topk_gpu.py tests/test_bi_gpu.py tests/test_sharded_gpu.py tests/test_maxsim_gpu.py -q
TAILN=20 run "perf f4: topk, bi" 400 python scripts/perf_f4.py topk bi | tee gpurun_out/perf_f4_b21.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
TAILN=3 run "launch list: perf_f4 bi" 300 ncu $M --log-file gpurun_out/launches_f4_b21.csv python scripts/perf_f4.py bi
