#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch10.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=12 run "pair check" 120 python scripts/pair_check.py check | tee gpurun_out/pair_check2.log
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=16 run "pair perf" 200 python scripts/pair_check.py perf | tee gpurun_out/pair_perf2.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard7.csv python scripts/loss_steps.py hard 3
TAILN=14 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles7.jsonl
TAILN=60 run "host profile" 300 python scripts/host_profile.py 200 | tee gpurun_out/host_profile.log
