#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_batch20.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=20 run "perf f4: head (graphs), topk (interleaved)" 400 python scripts/perf_f4.py head topk | tee gpurun_out/perf_f4_b20.jsonl
