#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests (new)"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "fused_decoder_tail or tensor_core" 2>&1 | tail -25
echo "== gpu tests" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | tee gpurun_out/test_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_iter.json | cut -c1-2500
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -45 gpurun_out/launch_summary.txt
echo "== grad parity (small goldens)"; timeout 600 python scripts/grad_parity_report.py --out gpurun_out/grad_parity.json 2>&1 | grep "mnist-small-mul\|taxibj" | cut -c1-1500
