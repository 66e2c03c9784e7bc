"""One traced pass of the persistent transformer-tower kernel at bench size: per-item
timestamps (SVDD_TOWER_TRACE) -> gpurun_out/tower_trace.csv, plus a summary of where an item's
time goes.  Tuning aid.
    python tools/tower_trace.py [--B 128] [--M 10]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import bench  # noqa: E402
from svdd_b200 import value_nets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--B', type=int, default=bench.B_PER_GPU)
ap.add_argument('--M', type=int, default=bench.M)
ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'tower_trace.csv'))
args = ap.parse_args()
dev = torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
cand = torch.randint(0, 4, (args.M * args.B, bench.L), device=dev, dtype=torch.uint8)
scorer = value_nets.packed_scorer(emb, head)
for _ in range(3):
  scorer.score(cand)
torch.cuda.synchronize()
os.makedirs(os.path.dirname(args.out), exist_ok=True)
if os.path.exists(args.out):
  os.unlink(args.out)
os.environ['SVDD_TOWER_TRACE'] = args.out
scorer.score(cand)
torch.cuda.synchronize()
del os.environ['SVDD_TOWER_TRACE']

rows = [l.strip().split(',') for l in open(args.out) if not l.startswith('#')]
rows = [[int(x) for x in r] for r in rows]
names = ['LN1', 'QKV', 'ATTN', 'OUT', 'LN2', 'FF1', 'FF2', 'BNACT']
end = max(max(r[6:12]) for r in rows)
print(f'kernel span {end / 1e3:.1f} us, {len(rows)} items')
import statistics as st
for q in range(8):
  rs = [r for r in rows if r[2] == q]
  if not rs:
    continue
  def col(a, b):
    v = [r[6 + b] - r[6 + a] for r in rs if r[6 + a] >= 0 and r[6 + b] >= 0]
    return f'{st.mean(v) / 1e3:6.2f}' if v else '   n/a'
  print(f'{names[q]:6s} n={len(rs):5d}  dep-wait {col(0, 1)}  dep->acc-done {col(1, 3)}  mma-start->acc-done {col(2, 3)}  '
        f'acc-done->epi-done {col(3, 4)}  epi-done->published {col(4, 5)}  dep->published {col(1, 5)} us')
# per block chain of row tile 0: when each phase completed
for j in (0, 5):
  line = []
  for q in range(7):
    rs = [r for r in rows if r[1] == j and r[2] == q and r[3] == 0]
    if rs:
      line.append(f'{names[q]}:{max(r[11] for r in rs) / 1e3:.1f}')
  print(f'block {j} row tile 0 published at (us):', ' '.join(line))
