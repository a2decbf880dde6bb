#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch12.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=30 run "stress" 300 python scripts/stress_loss.py 300 | tee gpurun_out/stress.log
if ! grep -q "STRESS OK" gpurun_out/stress.log; then
  TAILN=60 run "stress under compute-sanitizer" 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/stress_loss.py 3 | tee gpurun_out/stress_sanitizer.log
fi
run "suite" 900 python -m pytest tests -m gpu -q -x
TAILN=10 run "pair/issuer perf" 200 python scripts/pair_check.py perf | tee gpurun_out/pair_perf3.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard9.csv python scripts/loss_steps.py hard 3
TAILN=14 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles9.jsonl
TAILN=50 run "host profile" 300 python scripts/host_profile.py 200 | head -4 | tee gpurun_out/host_profile2.log
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b12.json
