#!/bin/bash
# 2-GPU session:   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_batch22.sh'
mkdir -p gpurun_out
N=${N:-$(nvidia-smi -L | wc -l)}
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-15}; echo "-- exit ${PIPESTATUS[0]}"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TAILN=12 run "top-k (new insertion) + 2-rank tests (sharded scorer, fused gather x50, training exchange)" 600 python -m pytest tests/test_topk_gpu.py tests/test_sharded_gpu.py tests/test_exchange_gpu.py -q
TAILN=6 run "cfg4 x$N (local top-10 fused into the kernel)" 420 $TR scripts/run_cfg4.py --docs-per-gpu ${DOCS:-12500} | tee gpurun_out/cfg4_topk_n$N.json
TAILN=4 run "bench x$N" 420 $TR bench.py --gpus $N --steps 100 --warmup 5 | tee gpurun_out/bench_b22_n$N.json
