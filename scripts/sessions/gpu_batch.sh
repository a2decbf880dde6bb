#!/bin/bash
# One GPU session: parity suite (with a fallback pass that switches the new launch / epilogue options off if the default
# pass fails, to localise a failure), then timings.   gpurun --timeout 1700 -- 'bash scripts/gpu_batch.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-12}; echo "-- exit ${PIPESTATUS[0]}"; }
run "suite (defaults: pdl=1 boundary_mode=1)" 900 python -m pytest tests -m gpu -q -x --deselect tests/test_sharded_gpu.py::test_two_ranks_nccl
if [ "${PIPESTATUS[0]}" != "0" ]; then :; fi
TAILN=25 run "suite again with pdl=0,boundary_mode=0 (only informative if the first pass failed)" 900 env COLPALI_B200_OPTS=pdl=0,boundary_mode=0 python -m pytest tests -m gpu -q -x -k "maxsim or loss" 
TAILN=40 run "variants" 300 python scripts/perf_variants.py | tee gpurun_out/perf_variants.jsonl
TAILN=40 run "aux" 600 python scripts/perf_aux.py | tee gpurun_out/perf_aux.jsonl
TAILN=40 run "wide" 600 python scripts/perf_wide.py | tee gpurun_out/perf_wide.jsonl
run "smoke" 300 python __graft_entry__.py --smoke
TAILN=5 run "bench" 600 python bench.py --steps 100 --warmup 5 | tee gpurun_out/bench_n1.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | tail -1
