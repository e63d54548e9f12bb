#!/bin/bash
# Gradient-parity distribution on the GPU, the tightened golden test, and full ncu captures of the two slowest tensor kernels.
mkdir -p gpurun_out/ncu4
timeout 300 python scripts/grad_parity_report.py --device cuda --out gpurun_out/grad_parity.json 2>&1 | grep -v -i warn | cut -c1-300
timeout 300 python -m pytest tests/test_step_gpu.py -q -m gpu -k golden --tb=short 2>&1 | tail -5
cap() {  # name, kernel regex, skip, count
  timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 -o /tmp/$1 -f python scripts/profile_step.py > gpurun_out/ncu4/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/ncu4/$1_details.txt 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu4/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/ncu4/$1_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu4/$1.log
}
cap conv_dec2_fprop 'tc_conv_kernel' 4 1
cap wgrad_dec2 'tc_wgrad_kernel' 0 1
du -sh gpurun_out/ncu4
