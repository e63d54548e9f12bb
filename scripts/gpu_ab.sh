#!/bin/bash
# A/B of environment switches: bash scripts/gpu_ab.sh "VAR=1" "VAR2=16" ...   (first run = defaults)
mkdir -p gpurun_out
run() { env $1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', {k:round(d[k],3) for k in ('value','ms_per_step')}, 'roof', round(d['roofline']['achieved'],1), 'bn_ms', round(d.get('roofline_hbm',{}).get('ms_per_step',0),3))"; }
run "X=0"
for v in "$@"; do run "$v"; done
