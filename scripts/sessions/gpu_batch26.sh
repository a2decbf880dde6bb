#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch26.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=25 run "losses (per-group fused loss tail)" 600 python -m pytest tests/test_loss_gpu.py tests/test_reference_replay_gpu.py -q
TAILN=8 run "stress" 300 python scripts/stress_loss.py 300 | tee gpurun_out/stress5.log
TAILN=8 run "loss step, CUDA graphs" 300 python scripts/loss_graph_time.py | tee gpurun_out/loss_graph_time_b26.jsonl
TAILN=6 run "suite" 900 python -m pytest tests -m gpu -q
TAILN=3 run "smoke" 300 python __graft_entry__.py --smoke
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b26.json
TAILN=4 run "perf f4: topk" 300 python scripts/perf_f4.py topk | tee gpurun_out/perf_f4_b26.jsonl
