#!/bin/bash
mkdir -p gpurun_out
echo "== tc kernels" ; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tensor_core" --tb=short 2>&1 | tail -60 | tee gpurun_out/test_tc.log
echo "== all kernels" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=line 2>&1 | tail -15 | tee gpurun_out/test_kernels.log
echo "== step" ; timeout 1200 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short 2>&1 | tail -40 | tee gpurun_out/test_step.log
echo "== bench bf16 graph" ; timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_bf16_graph.log
