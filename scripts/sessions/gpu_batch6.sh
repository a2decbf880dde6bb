#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_batch6.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard4.csv python scripts/loss_steps.py hard 3
TAILN=3 run "aux (loss lines)" 300 python scripts/perf_aux.py | head -2 | tee gpurun_out/perf_aux4.jsonl
TAILN=12 run "job cycles" 300 python scripts/job_cycles.py | tail -1 | tee gpurun_out/job_cycles4.jsonl
