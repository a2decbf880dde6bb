#!/bin/bash
# Multi-GPU session (N = number of visible GPUs, 2 by default):   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_batch2.sh'
mkdir -p gpurun_out
N=${N:-$(nvidia-smi -L | wc -l)}
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-15}; echo "-- exit ${PIPESTATUS[0]}"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TAILN=30 run "2-rank tests (sharded scorer, fused gather x50, training exchange)" 600 python -m pytest tests/test_sharded_gpu.py tests/test_exchange_gpu.py -q -x
TAILN=6 run "cfg4 x$N" 420 $TR scripts/run_cfg4.py --docs-per-gpu ${DOCS:-12500} | tee gpurun_out/cfg4_n$N.json
TAILN=6 run "exchange timing x$N" 300 $TR scripts/perf_exchange.py | tee gpurun_out/perf_exchange_n$N.json
TAILN=4 run "bench x$N" 420 $TR bench.py --gpus $N --steps 200 --warmup 5 | tee gpurun_out/bench_n$N.json
TAILN=4 run "bench x$N, NCCL gather (A/B)" 420 env COLPALI_B200_NCCL_GATHER=1 $TR bench.py --gpus $N --steps 200 --warmup 5 | tee gpurun_out/bench_n${N}_nccl.json
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv | tail -$N
