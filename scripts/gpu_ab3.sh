#!/bin/bash
# Fused-class kernel with the compile-time MMA schedule: opt-in parity test, A/B bench, launch list with it enabled.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "fused_class or layout or loss or pack" 2>&1 | tail -4
B="timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
echo "== bench default";       $B 2>&1 | tail -1 | tee gpurun_out/ab3_default.json | cut -c1-170
echo "== bench fused classes"; VARSEP_ENABLE_FUSED_CLASSES=1 $B 2>&1 | tail -1 | tee gpurun_out/ab3_fused.json | cut -c1-170
VARSEP_ENABLE_FUSED_CLASSES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ab3.csv python scripts/profile_step.py > gpurun_out/prof_ab3.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ab3.csv > gpurun_out/launch_summary_ab3.txt; grep -E "convT4|tc_conv_kernel|launches," gpurun_out/launch_summary_ab3.txt | cut -c1-150
