#!/bin/bash
# 2 GPUs: hardware data-parallel test + bench at N=1 and N=2 (same box, back to back)
mkdir -p gpurun_out
echo "== nccl test"; timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -q -m gpu --tb=short 2>&1 | tail -15
echo "== N=1"; timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>&1 | tail -1 | tee gpurun_out/scale_n1.json | cut -c1-200
echo "== N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -3 | tee gpurun_out/scale_n2.json | cut -c1-200
