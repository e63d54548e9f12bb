#!/bin/bash
# Fused-class transposed convolution: parity (kernel + step), A/B bench, launch list.
mkdir -p gpurun_out
echo "== kernel tests" ; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "conv or pack" 2>&1 | tail -8
echo "== step tests" ; timeout 300 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5
B="timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
echo "== bench new";               $B 2>&1 | tail -1 | tee gpurun_out/ab2_new.json | cut -c1-170
echo "== bench no fused classes";  VARSEP_DISABLE_FUSED_CLASSES=1 $B 2>&1 | tail -1 | tee gpurun_out/ab2_nofused.json | cut -c1-170
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ab2.csv python scripts/profile_step.py > gpurun_out/prof_ab2.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ab2.csv > gpurun_out/launch_summary_ab2.txt; head -48 gpurun_out/launch_summary_ab2.txt | cut -c1-150
