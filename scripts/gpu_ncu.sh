#!/bin/bash
mkdir -p gpurun_out
echo "== full bench (with cpu baseline)"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_full.json; cut -c1-300 gpurun_out/bench_full.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
echo "== ncu full: tc_conv_kernel"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -s 10 -c 3 -o gpurun_out/prof_tc_conv -f python scripts/profile_step.py > gpurun_out/ncu_tc_conv.log 2>&1; tail -2 gpurun_out/ncu_tc_conv.log
echo "== ncu full: tc_wgrad_kernel"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_wgrad_kernel -s 4 -c 2 -o gpurun_out/prof_tc_wgrad -f python scripts/profile_step.py > gpurun_out/ncu_tc_wgrad.log 2>&1; tail -2 gpurun_out/ncu_tc_wgrad.log
ls -la gpurun_out/*.ncu-rep
