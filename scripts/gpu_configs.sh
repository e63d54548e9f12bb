#!/bin/bash
# every BASELINE configuration at full size through bench.py (short), bf16 and fp32
mkdir -p gpurun_out
for cfg in wave taxibj chairs sst; do
  echo "== $cfg bf16"; timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${cfg}_bf16.log | cut -c1-330
done
echo "== mnist fp32"; timeout 600 python bench.py --config mnist --dtype fp32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_mnist_fp32.log | cut -c1-330
