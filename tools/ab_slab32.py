"""A/B of the 32-column slab protocol of EPI_PAIR / EPI_POOL2 (SVDD_SLAB32, read per call) on the c2
value pass (1280 candidates x L = 200, the bench network): bit-identical scores, graph-replay time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
from svdd_b200 import value_nets  # noqa: E402

dev = torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
cand = torch.randint(0, 4, (1280, 200), device=dev, dtype=torch.uint8)
scorer = value_nets.packed_scorer(emb, head)
outs = {}
for mode in ('0', '1', '0', '1'):
  os.environ['SVDD_SLAB32'] = mode
  out = torch.empty(1280, device=dev)
  for _ in range(2):
    scorer.score(cand, out=out)
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    scorer.score(cand, out=out)
  ts = []
  for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  outs[mode] = out.clone()
  print(f'SVDD_SLAB32={mode}: value pass (graph) min {min(ts):.4f} ms median {sorted(ts)[5]:.4f} ms')
print('bit-identical:', torch.equal(outs['0'], outs['1']), 'finite:', bool(torch.isfinite(outs['1']).all()))
