#!/bin/bash
# whole GPU suite + the bench lines of the other configurations and of the long-horizon rollout
mkdir -p gpurun_out
echo "== gpu tests" ; timeout 1800 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -25 | tee gpurun_out/test_gpu.log
for c in wave taxibj chairs sst; do
  echo "== bench $c"; timeout 900 python bench.py --config $c --steps 10 --warmup 3 --cpu-budget 8 --no-gpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$c.json | cut -c1-400
done
echo "== rollout"; timeout 600 python bench.py --mode rollout --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_rollout.json | cut -c1-1500
