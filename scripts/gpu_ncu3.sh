#!/bin/bash
mkdir -p gpurun_out/ncu3
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 -o /tmp/$1 -f python scripts/profile_step.py > gpurun_out/ncu3/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/ncu3/$1_details.txt 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/ncu3/$1_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu3/$1.log
}
cap im2col_fwd_k2 im2col_fwd_kernel 2 1
cap im2col_fwd_k4 im2col_fwd_kernel 0 1
du -sh gpurun_out/ncu3
