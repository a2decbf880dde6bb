#!/bin/bash
# Round-2 first GPU session for the drafts on branch r2-drafts (see DRAFTS.md).  Every stage runs under its own timeout
# so that a hung draft kernel (the device-side watchdog in mbar_wait traps after 4 s) cannot eat the session.
#   gpurun --timeout 1500 -- 'bash scripts/r2_validate.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -8; echo "-- exit ${PIPESTATUS[0]}"; }
# 0. the validated product still passes (regression gate for everything the drafts touched: cabi.cu, losses.py, head.py)
run "validated suite" 900 python -m pytest tests -m gpu -x -q -k "not wide"
# 1. drafts, smallest first; -x inside each file so the first failure is reported with its traceback
run "wide head, cluster 1" 300 python -m pytest tests/test_head_gpu.py -q -x -k "wide and not cluster2"
run "wide head, cluster 2" 300 python -m pytest tests/test_head_gpu.py -q -x -k "wide and cluster2"
run "K-pipelined scorer" 300 python -m pytest tests/test_maxsim_gpu.py -q -x -k wide
run "wide loss + backward" 300 python -m pytest tests/test_loss_gpu.py -q -x -k wide
# 2. timings
run "perf" 600 python scripts/r2_perf_wide.py | tee gpurun_out/r2_perf_wide.jsonl
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | tail -1
