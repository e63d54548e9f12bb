#!/bin/bash
# data-parallel bench on N GPUs of one node (torchrun, one rank per GPU, NCCL)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for variant in "" "--no-overlap"; do
  echo "== N=$N variant='$variant'"
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $variant > gpurun_out/bench_n${N}${variant// /_}.log 2>&1
  tail -1 gpurun_out/bench_n${N}${variant// /_}.log | cut -c1-260
done
