#!/bin/bash
# ncu --set full of every launch of one kernel family in one eager step: bash scripts/gpu_ncu_kernel.sh <regex> <name> [count]
R=$1; NAME=$2; C=${3:-12}
O=gpurun_out/ncu; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$R" -c $C -o /tmp/$NAME -f python scripts/profile_step.py > $O/$NAME.log 2>&1; tail -1 $O/$NAME.log
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > /tmp/$NAME.csv 2>/dev/null
python scripts/summarize_ncu_raw.py /tmp/$NAME.csv > $O/${NAME}_per_launch.txt; cat $O/${NAME}_per_launch.txt
ncu -i /tmp/$NAME.ncu-rep --page details --launch-skip 0 --launch-count 1 > $O/${NAME}_details_first.txt 2>/dev/null
cp /tmp/$NAME.ncu-rep $O/ 2>/dev/null
