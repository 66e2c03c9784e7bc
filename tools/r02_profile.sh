#!/bin/bash
# Round-2 evidence pass on the GPU box (run under gpurun):
#   * full -m gpu test suite, bench.py (N = 1) and its reference arm
#   * ncu launch list (time, tensor pipe, DRAM bytes) of one eager c2 reverse step, one RNA
#     SVDD-PM step (c5-shaped, B = 256) and one DiT forward
#   * ncu --set full of the tower kernel, the fused denoiser at L = 50 (one item per CTA and
#     den_short_kernel), the GRU recurrence and the DiT attention kernel; raw pages as CSV
#   gpurun --timeout 2400 -- 'bash tools/r02_profile.sh r02'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
M1=gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
python -m pytest tests -q -m gpu 2>&1 | tail -15 > $OUT/${TAG}_pytest.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
timeout 300 python tools/gemm_table.py > $OUT/${TAG}_gemm_table.txt 2>&1
timeout 600 ncu --metrics $M1 --clock-control none --nvtx --nvtx-include "measured/" \
    --csv --log-file $OUT/${TAG}_launches_c2_step.csv python tools/profile_step.py --steps 1 > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics $M1 --clock-control none --nvtx --nvtx-include "measured/" \
    --csv --log-file $OUT/${TAG}_launches_c5_step.csv python tools/profile_step_rna.py --steps 1 > $OUT/${TAG}_launches_c5.log 2>&1
for spec in "tower:tower_kernel:profile_step.py:0" "den_single_l50:den_fused_kernel:profile_step_rna.py:1" "den_short:den_short_kernel:profile_step_rna.py:0" "gru_umma:cg_gru_umma_kernel:profile_step_rna.py --B 1024:0" "dit_attn:dit_attn_small_kernel:time_dit.py 64 200:1"; do
  name=${spec%%:*}; rest=${spec#*:}; re=${rest%%:*}; rest=${rest#*:}; script=${rest%%:*}; skip=${rest#*:}
  inc="--nvtx --nvtx-include measured/"
  if [ "$name" = "dit_attn" ]; then inc=""; export SVDD_TIME_DIT_NO_PROFILE=1; fi
  timeout 900 ncu --set full --clock-control none --import-source on $inc -k regex:"$re" -s $skip -c 1 \
      -o $OUT/${TAG}_${name}_full -f python tools/$script > $OUT/${TAG}_${name}_full.log 2>&1
  ncu -i $OUT/${TAG}_${name}_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_${name}_raw.csv 2>/dev/null
  rm -f $OUT/${TAG}_${name}_full.ncu-rep
done
# the residual 1x1 conv / difference pooling at stage 0 (2nd / 3rd gemm2 launch of the step): raw + top stalls of the source page
for spec in "pair:1" "pool2:2"; do
  name=${spec%%:*}; skip=${spec#*:}
  timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include measured/ -k regex:gemm2_kernel -s $skip -c 1 \
      -o $OUT/${TAG}_${name}_full -f python tools/profile_step.py --steps 1 > $OUT/${TAG}_${name}_full.log 2>&1
  ncu -i $OUT/${TAG}_${name}_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_${name}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}_full.ncu-rep --page source --csv > $OUT/${TAG}_ncu_full_${name}_source.csv 2>/dev/null
  python tools/ncu_top_stalls.py $OUT/${TAG}_ncu_full_${name}_source.csv 0 25 > $OUT/${TAG}_ncu_${name}_top_stalls.txt 2>&1
  rm -f $OUT/${TAG}_${name}_full.ncu-rep $OUT/${TAG}_ncu_full_${name}_source.csv
done
ls -la $OUT | tail -25
tail -5 $OUT/${TAG}_pytest.log
head -c 1500 $OUT/${TAG}_bench.json
