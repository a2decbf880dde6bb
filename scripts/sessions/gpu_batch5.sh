#!/bin/bash
# Validation + timing of the round's last kernel changes (1 GPU):  gpurun --timeout 1200 -- 'bash scripts/gpu_batch5.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=12 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles3.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss smooth" 300 ncu $M --log-file gpurun_out/launches_loss_smooth3.csv python scripts/loss_steps.py smooth 3
TAILN=16 run "wide" 600 python scripts/perf_wide.py | tee gpurun_out/perf_wide3.jsonl
