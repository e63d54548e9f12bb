#!/bin/bash
# ncu evidence, kept small: per-launch section metrics of every tensor-core launch as CSV, and one
# `--set full --import-source on` capture of a single representative launch per kernel.
mkdir -p gpurun_out
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --section SchedulerStats"
timeout 1200 ncu $SECT --clock-control none --profile-from-start off -k regex:tc_ -c 120 -o /tmp/tc_all -f python scripts/profile_step.py > gpurun_out/ncu_tc_all.log 2>&1
ncu -i /tmp/tc_all.ncu-rep --page raw --csv > gpurun_out/tc_all_raw.csv 2>/dev/null
ls -la /tmp/tc_all.ncu-rep gpurun_out/tc_all_raw.csv
# decoder conv.2 forward = 11th tc_conv launch of the step (Es 4, Et 4, decoder conv.0, conv.1, conv.2)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -s 10 -c 1 -o gpurun_out/prof_tc_conv_dec2 -f python scripts/profile_step.py > gpurun_out/ncu_tc_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -s 8 -c 1 -o gpurun_out/prof_tc_conv_dec0 -f python scripts/profile_step.py >> gpurun_out/ncu_tc_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_wgrad_kernel -s 0 -c 1 -o gpurun_out/prof_tc_wgrad_dec2 -f python scripts/profile_step.py > gpurun_out/ncu_tc_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
