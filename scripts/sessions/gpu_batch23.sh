#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_batch23.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=6 run "suite" 900 python -m pytest tests -m gpu -q
TAILN=3 run "smoke" 300 python __graft_entry__.py --smoke
TAILN=3 run "bench" 300 python bench.py --steps 200 --warmup 20 | tee gpurun_out/bench_b23.json
TAILN=3 run "bench --impl reference" 300 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_b23_ref.json
NCU="ncu --set full --clock-control none --import-source on -f"
TAILN=3 run "ncu: dense 128-tile" 300 $NCU -k regex:dense_tile_kernel -s 1 -c 1 -o gpurun_out/prof_dense128 python scripts/one_shot.py dense
TAILN=3 run "ncu: maxsim + fused top-k" 300 $NCU -k regex:maxsim_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_maxsim_topk python scripts/one_shot.py topk
ls -la gpurun_out/*.ncu-rep
