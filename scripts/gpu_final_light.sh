#!/bin/bash
# Light validation: parity suite, default bench line, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -2 | tee gpurun_out/test_gpu.log
timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json; cut -c1-300 gpurun_out/bench_full.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt; head -3 gpurun_out/launch_summary.txt
