#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_batch24.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=15 run "bi + losses" 600 python -m pytest tests/test_bi_gpu.py tests/test_loss_gpu.py -q
TAILN=8 run "perf f4: bi" 300 python scripts/perf_f4.py bi | tee gpurun_out/perf_f4_b24.jsonl
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --csv"
TAILN=3 run "ncu: dense 128-tile traffic" 300 ncu $M -k regex:dense_tile_kernel -s 1 -c 1 --log-file gpurun_out/dense128_traffic_b24.csv python scripts/one_shot.py dense
cat gpurun_out/dense128_traffic_b24.csv | tail -8
