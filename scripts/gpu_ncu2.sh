#!/bin/bash
# full-set ncu captures of the epilogue-/bandwidth-bound kernels; raw metric tables come back as CSV
mkdir -p gpurun_out/ncu2
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 -o /tmp/$1 -f python scripts/profile_step.py > gpurun_out/ncu2/$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu2/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/ncu2/$1_details.txt 2>/dev/null
  tail -1 gpurun_out/ncu2/$1.log
}
cap im2col_fwd im2col_fwd_kernel 2 1
cap tc_conv64 "tc_conv_kernel<64" 0 6
cap bn_fwd bn_act_fwd_col 9 1
cap bn_bwd bn_bwd_apply_col 0 1
cap im2col_wgrad im2col_wgrad_kernel 0 1
du -sh gpurun_out/ncu2
