#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch15.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=8 run "stress" 300 python scripts/stress_loss.py 400 | tee gpurun_out/stress4.log
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard12.csv python scripts/loss_steps.py hard 3
TAILN=14 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles10.jsonl
TAILN=3 run "aux" 400 python scripts/perf_aux.py | head -2 | tee gpurun_out/perf_aux7.jsonl
