#!/bin/bash
# One iteration: full GPU parity suite, bench line, launch list, full ncu captures named on the command line
# (name:regex:skip triples), e.g.  bash scripts/gpu_iter.sh im2col_enc0:im2col_fwd_kernel:0
mkdir -p gpurun_out/ncu5
echo "== gpu tests" ; timeout 900 python -m pytest tests -q -m gpu --tb=short -x 2>&1 | tail -6
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/iter_bench.json | cut -c1-170
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_iter.csv python scripts/profile_step.py > gpurun_out/prof_iter.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_iter.csv > gpurun_out/launch_summary_iter.txt; head -48 gpurun_out/launch_summary_iter.txt
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:$regex" -s $skip -c 1 -o /tmp/$name -f python scripts/profile_step.py > gpurun_out/ncu5/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/ncu5/${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/ncu5/${name}_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu5/$name.log
done
