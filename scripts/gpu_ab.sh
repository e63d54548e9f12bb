#!/bin/bash
# Kernel + step parity for the new epilogue / weight-gradient variants, then A/B bench lines and a launch list.
mkdir -p gpurun_out
echo "== kernel tests" ; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "conv" 2>&1 | tail -8
echo "== step tests" ; timeout 600 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
echo "== bench new";            $B 2>&1 | tail -1 | tee gpurun_out/ab_new.json | cut -c1-170
echo "== bench no staged stats"; VARSEP_DISABLE_STAGED_STATS=1 $B 2>&1 | tail -1 | tee gpurun_out/ab_nostats.json | cut -c1-170
echo "== bench no wgrad taps";   VARSEP_DISABLE_WGRAD_TAPS=1 $B 2>&1 | tail -1 | tee gpurun_out/ab_notaps.json | cut -c1-170
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ab.csv python scripts/profile_step.py > gpurun_out/prof_ab.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_ab.csv > gpurun_out/launch_summary_ab.txt; head -44 gpurun_out/launch_summary_ab.txt
