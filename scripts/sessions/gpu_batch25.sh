#!/bin/bash
# gpurun --timeout 600 -- 'bash scripts/gpu_batch25.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=15 run "bi + a5 training shapes" 300 python -m pytest tests/test_bi_gpu.py tests/test_loss_gpu.py -q -k "dense or training_shapes"
TAILN=3 run "dense raster A/B" 200 python scripts/dense_raster_ab.py | tee gpurun_out/dense_raster_ab.json
