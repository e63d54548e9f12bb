#!/bin/bash
mkdir -p gpurun_out
echo "== kernels" ; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=line 2>&1 | tail -25 | tee gpurun_out/test_kernels.log
echo "== step" ; timeout 1200 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short 2>&1 | tail -15 | tee gpurun_out/test_step.log
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -30 gpurun_out/launch_summary.txt
echo "== bench bf16 graph" ; timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_graph.log | cut -c1-200
if false; then
echo "== ncu full: all tc_conv launches of one step"
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -c 100 -o gpurun_out/prof_tc_conv -f python scripts/profile_step.py > gpurun_out/ncu_tc_conv.log 2>&1; tail -2 gpurun_out/ncu_tc_conv.log
fi
