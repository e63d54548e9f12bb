#!/bin/bash
# Iteration: full GPU parity suite, bench, launch list, plus a full ncu capture of the opt-in fused-class kernel.
mkdir -p gpurun_out/ncu6
echo "== gpu tests" ; timeout 900 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -6
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/iter_bench.json | cut -c1-170
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_iter.csv python scripts/profile_step.py > gpurun_out/prof_iter.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_iter.csv > gpurun_out/launch_summary_iter.txt; head -30 gpurun_out/launch_summary_iter.txt
name=convT4_dec2
VARSEP_ENABLE_FUSED_CLASSES=1 timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:tc_convT4_kernel" -s 2 -c 1 -o /tmp/$name -f python scripts/profile_step.py > gpurun_out/ncu6/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/ncu6/${name}_details.txt 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/ncu6/${name}_source.csv 2>/dev/null
tail -1 gpurun_out/ncu6/$name.log
