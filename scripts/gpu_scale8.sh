#!/bin/bash
# 8-GPU box: N=1 and N=8 (peer-memory exchange, bf16 and fp32 wire) back to back
mkdir -p gpurun_out/r02
run() { n=$1; shift; tag=$1; shift; if [ $n = 1 ]; then python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 30 --warmup 5 "$@"; fi 2>&1 | grep '^{' | tail -1 > gpurun_out/r02/scale_$tag.json; python -c "
import json; d=json.load(open('gpurun_out/r02/scale_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],3), d['run'].get('gradient_exchange'))"; }
run 1 n1
run 8 n8_peer_bf16
run 8 n8_peer_fp32 --wire fp32
run 4 n4_peer_bf16
