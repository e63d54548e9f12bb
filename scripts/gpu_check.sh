#!/bin/bash
# One GPU session: kernel tests, step tests, smoke, short benches.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels" ; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -40 | tee gpurun_out/test_kernels.log
echo "== step" ; timeout 1200 python -m pytest tests/test_step_gpu.py -q -m gpu --tb=short 2>&1 | tail -60 | tee gpurun_out/test_step.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench fp32 eager" ; timeout 600 python bench.py --steps 3 --warmup 3 --dtype fp32 --no-graph --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_fp32_eager.log
echo "== bench bf16 eager" ; timeout 600 python bench.py --steps 3 --warmup 3 --dtype bf16 --no-graph --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_bf16_eager.log
echo "== bench bf16 graph" ; timeout 900 python bench.py --steps 5 --warmup 3 --dtype bf16 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_bf16_graph.log
