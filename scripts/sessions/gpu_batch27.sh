#!/bin/bash
# 2-GPU session:   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_batch27.sh'
mkdir -p gpurun_out
N=${N:-$(nvidia-smi -L | wc -l)}
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-15}; echo "-- exit ${PIPESTATUS[0]}"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TAILN=8 run "2-rank tests after the per-group loss tail (training exchange: in-kernel wait + fused loss; sharded scorer)" 600 python -m pytest tests/test_exchange_gpu.py tests/test_sharded_gpu.py -q
TAILN=6 run "exchange timing x$N" 300 $TR scripts/perf_exchange.py | tee gpurun_out/perf_exchange_b27_n$N.json
