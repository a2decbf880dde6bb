#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch11.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard8.csv python scripts/loss_steps.py hard 3
TAILN=14 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles8.jsonl
TAILN=6 run "host profile" 300 python scripts/host_profile.py 200 | head -3 | tee gpurun_out/host_profile2.log
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b11.json
