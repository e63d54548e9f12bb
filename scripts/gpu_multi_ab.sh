#!/bin/bash
# N GPUs (first argument): A/B of the gradient-exchange settings, one short bench each
N=$1; shift
run() { echo "== $*"; env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 "$@" 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:round(d[k],3) for k in ('value','ms_per_step')})"; }
ENVV="X=0" run
ENVV="X=0" run --wire fp32
ENVV="X=0" run --no-overlap
ENVV="X=0" run --early-blocks 12
