#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch14.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=8 run "stress" 300 python scripts/stress_loss.py 400 | tee gpurun_out/stress3.log
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard11.csv python scripts/loss_steps.py hard 3
run "launch list: loss smooth" 300 ncu $M --log-file gpurun_out/launches_loss_smooth4.csv python scripts/loss_steps.py smooth 3
TAILN=4 run "host profile" 300 python scripts/host_profile.py 200 | head -3 | tee gpurun_out/host_profile3.log
TAILN=6 run "aux" 400 python scripts/perf_aux.py | tee gpurun_out/perf_aux6.jsonl
