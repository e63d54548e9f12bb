#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}.log 2>&1
tail -1 gpurun_out/bench_n${N}.log | cut -c1-300
