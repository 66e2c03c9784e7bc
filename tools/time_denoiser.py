"""CUDA-event timing of the fused denoiser: c2 size (B=128, L=200) and c5-shard size (51 200 x L=50)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers  # noqa: E402

dev = torch.device('cuda:0')
for L, n in ((200, 128), (200, 1280), (50, 1024), (50, 51200), (50, 101)):
  m = helpers.build_denoiser(44, L).to(dev)
  den = m.packed()
  x = helpers.random_tokens(n, L, 3, 0.5).to(dev).to(torch.uint8)
  out = torch.empty((n, L, 5), device=dev)
  for _ in range(3):
    den.forward(x, 0.0, out=out)
  torch.cuda.synchronize()
  ts = []
  for _ in range(7):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); den.forward(x, 0.0, out=out); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  print(f'L={L:4d} n={n:6d}  {sorted(ts)[3]:8.3f} ms')
