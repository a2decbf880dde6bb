#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_batch19.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=30 run "head + replay" 600 python -m pytest tests/test_head_gpu.py tests/test_reference_replay_gpu.py -q
TAILN=20 run "perf f4" 400 python scripts/perf_f4.py | tee gpurun_out/perf_f4.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
TAILN=3 run "launch list: perf_f4 bi" 300 ncu $M --log-file gpurun_out/launches_f4.csv python scripts/perf_f4.py bi
