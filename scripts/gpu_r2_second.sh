#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests" ; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -15 | tee gpurun_out/test_gpu.log
echo "== fullsize parity" ; timeout 900 python scripts/fullsize_parity_report.py --out gpurun_out/fullsize_parity.json 2>&1 | tail -8 | cut -c1-2500
echo "== grad parity (small goldens)"; timeout 600 python scripts/grad_parity_report.py --out gpurun_out/grad_parity.json 2>&1 | tail -10 | cut -c1-400
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_full.json | cut -c1-3000
