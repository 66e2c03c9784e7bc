#!/bin/bash
# Profiling pass run on the GPU box (under gpurun): per-GEMM warm table, ncu launch list of
# one eager reverse step of the bench workload, and one `--set full` capture of a few GEMM
# launches.  Outputs land in gpurun_out/ with the given tag (gpurun brings back <= 64 MiB, so
# the raw/source pages are exported to CSV on the box and an oversized .ncu-rep is dropped).
#   gpurun --timeout 1200 -- 'bash tools/gpu_profile.sh r01b [full|list] [n_full] [skip_full]'
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
M1=gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 python tools/gemm_table.py > $OUT/${TAG}_gemm_table.txt 2>&1
timeout 600 ncu --metrics $M1 --clock-control none --nvtx --nvtx-include "measured/" \
    --csv --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py --steps 1 > $OUT/${TAG}_launches.log 2>&1
if [ "${2:-full}" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measured/" \
      -k regex:'gemm2_kernel|conv_gemm_kernel' -s ${4:-0} -c ${3:-10} -o $OUT/${TAG}_gemm_full -f \
      python tools/profile_step.py --steps 1 > $OUT/${TAG}_full.log 2>&1
  ncu -i $OUT/${TAG}_gemm_full.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_full_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_gemm_full.ncu-rep --page details --csv > $OUT/${TAG}_gemm_full_details.csv 2>/dev/null
  ncu -i $OUT/${TAG}_gemm_full.ncu-rep --page source --csv > $OUT/${TAG}_gemm_full_source.csv 2>/dev/null
  SZ=$(stat -c %s $OUT/${TAG}_gemm_full.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -gt 30000000 ]; then rm -f $OUT/${TAG}_gemm_full.ncu-rep; fi
  gzip -f $OUT/${TAG}_gemm_full_source.csv
fi
du -sh $OUT; ls -la $OUT
tail -5 $OUT/${TAG}_gemm_table.txt
