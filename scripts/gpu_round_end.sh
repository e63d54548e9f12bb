#!/bin/bash
# Round-end validation of HEAD in one GPU session: parity tests, smoke, launch list, DRAM traffic of the tensor-core
# convolutions (-> profiles/roofline_traffic.json, read by bench.py), default bench line, reference arm.
mkdir -p gpurun_out
echo "== gpu tests" ; timeout 1200 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -6 | tee gpurun_out/test_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -12 gpurun_out/launch_summary.txt
echo "== traffic + bench default + bench reference"
bash scripts/gpu_traffic.sh 2>&1 | cut -c1-400
