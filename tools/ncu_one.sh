#!/bin/bash
# One `ncu --set full` capture of a kernel, raw + source pages as CSV under gpurun_out/:
#   gpurun --timeout 900 -- 'bash tools/ncu_one.sh r02 den_short den_short_kernel 0 profile_step_rna.py'
TAG=$1; NAME=$2; RE=$3; SKIP=$4; shift 4
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include measured/ -k regex:"$RE" -s $SKIP -c 1 \
    -o $OUT/${TAG}_${NAME}_full -f python tools/"$@" > $OUT/${TAG}_${NAME}_full.log 2>&1
ncu -i $OUT/${TAG}_${NAME}_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_${NAME}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_${NAME}_full.ncu-rep --page source --csv > $OUT/${TAG}_ncu_full_${NAME}_source.csv 2>/dev/null
rm -f $OUT/${TAG}_${NAME}_full.ncu-rep
ls -la $OUT | grep ${TAG}_ncu_full_${NAME}
