#!/bin/bash
# quick iteration: selected kernel tests, one bench line, launch list
mkdir -p gpurun_out
echo "== kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "${1:-fused_decoder_tail}" 2>&1 | tail -8
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_iter.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'hbm', d.get('roofline_hbm',{}).get('ms_per_step'))"
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -${2:-50} gpurun_out/launch_summary.txt
