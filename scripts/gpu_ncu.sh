#!/bin/bash
# ncu evidence, kept small: SpeedOfLight/Memory sections of EVERY launch of one step as CSV, plus one
# `--set full --import-source on` capture of a representative launch of each tensor-core kernel.
mkdir -p gpurun_out
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy"
timeout 1500 ncu $SECT --clock-control none --profile-from-start off -o /tmp/step_all -f python scripts/profile_step.py > gpurun_out/ncu_step_all.log 2>&1
ncu -i /tmp/step_all.ncu-rep --page raw --csv > /tmp/step_all_raw.csv 2>/dev/null
python scripts/summarize_ncu_raw.py /tmp/step_all_raw.csv > gpurun_out/ncu_step_all_summary.txt; head -5 gpurun_out/ncu_step_all_summary.txt
# decoder conv.2 forward = 11th tc_conv launch of the step (Es 4, Et 4, decoder conv.0, conv.1, conv.2); conv.0 = 9th
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -s 10 -c 1 -o gpurun_out/prof_tc_conv_dec2 -f python scripts/profile_step.py > gpurun_out/ncu_tc_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_conv_kernel -s 8 -c 1 -o gpurun_out/prof_tc_conv_dec0 -f python scripts/profile_step.py >> gpurun_out/ncu_tc_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_wgrad_kernel -s 0 -c 1 -o gpurun_out/prof_tc_wgrad_dec2 -f python scripts/profile_step.py > gpurun_out/ncu_tc_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
