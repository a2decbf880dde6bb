#!/bin/bash
# 8-GPU session:   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_batch8.sh'
mkdir -p gpurun_out
N=${N:-$(nvidia-smi -L | wc -l)}
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-8}; echo "-- exit ${PIPESTATUS[0]}"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TAILN=4 run "cfg4 x$N (128 q x $((12500 * N)) docs)" 420 $TR scripts/run_cfg4.py --docs-per-gpu 12500 | tee gpurun_out/cfg4_n$N.json
TAILN=4 run "cfg5 x$N (1000 pages)" 600 $TR scripts/run_cfg5.py --pages 1000 --batch 25 | tee gpurun_out/cfg5_n$N.json
TAILN=4 run "exchange timing x$N" 300 $TR scripts/perf_exchange.py | tee gpurun_out/perf_exchange_n$N.json
TAILN=3 run "bench x$N" 420 $TR bench.py --gpus $N --steps 200 --warmup 5 | tee gpurun_out/bench_n$N.json
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv | tail -$N
