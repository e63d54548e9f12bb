#!/bin/bash
# Round-2 first GPU session: whole GPU suite, smoke, full-size parity report, bench line (+ reference-gpu arm), launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -30 | tee gpurun_out/test_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== fullsize parity" ; timeout 900 python scripts/fullsize_parity_report.py --out gpurun_out/fullsize_parity.json 2>&1 | tail -8 | cut -c1-1500
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_full.json | cut -c1-3000
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -60 gpurun_out/launch_summary.txt
