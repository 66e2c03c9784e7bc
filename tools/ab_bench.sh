#!/bin/bash
# A/B helper: runs bench.py under several environment settings and prints one line each.
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1),'seq/s', round(d['ms_per_denoise_step'],3),'ms/step GEMM', round(d['roofline']['achieved'],1),'TFLOP/s', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
