#!/bin/bash
# Profiling session (1 GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_batch3.sh'
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout "$@" 2>&1 | tail -${TAILN:-12}; echo "-- exit ${PIPESTATUS[0]}"; }
TAILN=12 run "job cycles" 300 python scripts/job_cycles.py | tee gpurun_out/job_cycles.jsonl
TAILN=12 run "variants (interleaved)" 600 python scripts/perf_variants.py | tee gpurun_out/perf_variants.jsonl
M="--metrics gpu__time_duration.sum --clock-control none --csv"
run "launch list: loss hard" 300 ncu $M --log-file gpurun_out/launches_loss_hard.csv python scripts/loss_steps.py hard 3
run "launch list: loss smooth" 300 ncu $M --log-file gpurun_out/launches_loss_smooth.csv python scripts/loss_steps.py smooth 3
run "launch list: bench" 400 ncu $M -c 300 --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu
F="--set full --clock-control none --import-source on"
run "ncu full: maxsim" 600 ncu $F -k regex:maxsim_fwd_kernel -s 5 -c 2 -o gpurun_out/prof_maxsim python bench.py --steps 6 --warmup 3 --no-cpu
run "ncu full: loss path" 600 ncu $F -k "regex:maxsim_bwd|maxsim_fwd" -s 4 -c 3 -o gpurun_out/prof_loss python scripts/loss_steps.py hard 3
run "ncu full: smooth bwd" 600 ncu $F -k regex:smooth_bwd -s 2 -c 2 -o gpurun_out/prof_smooth python scripts/loss_steps.py smooth 2
run "ncu full: head" 600 ncu $F -k regex:head_ -c 2 -o gpurun_out/prof_head python -c "
import torch, colpali_b200 as cb
dev=torch.device('cuda:0')
for hidden, dim in ((1536,128),(2048,320)):
    h=torch.randn(275000,hidden,device=dev).bfloat16(); lin=torch.nn.Linear(hidden,dim).to(dev,torch.bfloat16)
    cb.fused_head(h,lin.weight,lin.bias,torch.ones(275000,dtype=torch.long,device=dev)); torch.cuda.synchronize()
"
for f in gpurun_out/prof_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
done
ls -la gpurun_out/
