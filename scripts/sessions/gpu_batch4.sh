#!/bin/bash
# Quick validation of the training-path kernels (1 GPU):  gpurun --timeout 900 -- 'bash scripts/gpu_batch4.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard2.csv python scripts/loss_steps.py hard 3
run "launch list: loss smooth" 300 ncu $M --log-file gpurun_out/launches_loss_smooth2.csv python scripts/loss_steps.py smooth 3
TAILN=10 run "aux" 600 python scripts/perf_aux.py | tee gpurun_out/perf_aux2.jsonl
TAILN=10 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles2.jsonl
TAILN=3 run "bench" 600 python bench.py --steps 100 --warmup 5 | tee gpurun_out/bench_n1_b.json
