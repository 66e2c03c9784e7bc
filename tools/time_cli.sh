#!/bin/bash
# Wall clock of the reference's CLI surface (the SVDD batches AND the
# gen_batch_num * sample_M baseline rollouts + final scoring inside controlled_decode), random-init
# networks of the named architectures.  Each run prints decode.py's own `timing:` line (phase
# seconds, one device synchronisation per phase) and the process wall clock incl. start-up.
#   gpurun --timeout 1200 -- 'bash tools/time_cli.sh > gpurun_out/r02_cli_timing.txt 2>&1'
run() {
  echo "\$ python $* --random_init"
  local t0=$(date +%s.%N)
  python "$@" --random_init --out_dir /tmp/svdd_cli_log 2>&1 | grep -E "^timing|^decoding|Error|error"
  local t1=$(date +%s.%N)
  echo "process wall clock $(python -c "print(round($t1 - $t0, 2))") s"
}
run decode.py --task dna --sample_M 10 --batch_size 128 --val_batch_num 2
run decode_tweedie.py --tweedie True --task dna --sample_M 10 --batch_size 128 --val_batch_num 2
run decode.py --task rna --reward_name MRL --sample_M 10 --batch_size 10 --val_batch_num 2
run decode_tweedie.py --tweedie True --task rna --reward_name MRL --sample_M 50 --batch_size 1024 --val_batch_num 1
