#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), aux timings, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== aux"; timeout 600 python scripts/perf_aux.py 2>&1 | grep what | tee gpurun_out/perf_aux.jsonl
if [ "$1" != "noncu" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-300
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_fwd -s 5 -c 2 -f -o gpurun_out/prof_maxsim python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
fi
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | tail -1
lscpu | grep -E "Model name|^CPU\(s\)"
