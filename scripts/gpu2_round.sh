#!/bin/bash
# 2-GPU session: sharded tests, cfg4 at reduced and full per-GPU size, N=2 bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/run_cfg4.py --docs-per-gpu 12500 --check-sequential 2>&1 | tail -1 | tee gpurun_out/cfg4_n2.json
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 2>&1 | tail -1 | tee gpurun_out/bench_n2.json | cut -c1-300
