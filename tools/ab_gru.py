"""A/B of the ConvGRU scoring pass: tcgen05 recurrence (SVDD_GRU_UMMA=1) vs mma.sync + projection GEMM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import synthetic, value_nets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 51200
dev = torch.device('cuda:0')
emb, head = synthetic.build_convgru_oracle()
emb, head = emb.to(dev), head.to(dev)
tok = synthetic.random_tokens(n, 50, 3, 0.3).to(dev).to(torch.uint8)
res = {}
for mode in ('0', '1', '0', '1'):
  os.environ[os.environ.get('AB_VAR', 'SVDD_GRU_UMMA')] = mode
  out = value_nets.score_tokens(emb, head, tok)
  torch.cuda.synchronize()
  ts = []
  for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = value_nets.score_tokens(emb, head, tok); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  res[mode] = out.clone()
  print(f'{os.environ.get("AB_VAR", "SVDD_GRU_UMMA")}={mode}: ConvGRU score of {n} x 50: min {min(ts):.3f} ms median {sorted(ts)[2]:.3f} ms')
print('max |d|', float((res['0'] - res['1']).abs().max()))
