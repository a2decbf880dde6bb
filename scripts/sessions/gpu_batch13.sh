#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch13.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=12 run "stress" 300 python scripts/stress_loss.py 400 | tee gpurun_out/stress2.log
TAILN=12 run "stress again" 300 python scripts/stress_loss.py 400 | tee -a gpurun_out/stress2.log
run "suite" 900 python -m pytest tests -m gpu -q -x
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard10.csv python scripts/loss_steps.py hard 3
TAILN=4 run "ncu full: argmax forward" 600 ncu --set full --import-source on --clock-control none -k regex:maxsim_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_argmax python scripts/loss_steps.py hard 3
