#!/bin/bash
# One `ncu --set full` capture of the launches matching a kernel-name regex inside one eager
# reverse step of the bench workload; raw / details / source pages exported to CSV.
#   gpurun --timeout 900 -- 'bash tools/gpu_profile_kernel.sh <tag> <kernel regex> [count] [skip]'
TAG=${1:-prof}; RE=${2:-gemm2_kernel}; CNT=${3:-1}; SKIP=${4:-0}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measured/" \
    -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/${TAG}_full -f \
    python tools/profile_step.py --steps 1 > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page details --csv > $OUT/${TAG}_full_details.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > $OUT/${TAG}_full_source.csv 2>/dev/null
gzip -f $OUT/${TAG}_full_source.csv
SZ=$(stat -c %s $OUT/${TAG}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f $OUT/${TAG}_full.ncu-rep; fi
tail -3 $OUT/${TAG}_full.log
