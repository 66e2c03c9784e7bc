"""Stage 2 (svdd_subs_sample) and stage 4 (svdd_select_gather) at BASELINE config-4 size for ncu:
    ncu --set full --clock-control none -k regex:'subs_sample|select_gather' -s 2 -c 2 -o out python tools/profile_sampling.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import _lib  # noqa: E402

B, L, M = int(os.environ.get('B', 4096)), 200, 20
dev = torch.device('cuda:0')
lg = torch.randn(B, L, 5, device=dev)
xs = torch.full((B, L), 4, dtype=torch.int64, device=dev)
U = torch.rand(M, B, L, 5, device=dev)
sc = torch.randn(M, B, device=dev)
cand = _lib.subs_sample(lg, xs, M, 0.5, 0.49, U=U)
x = _lib.select_gather(sc, cand)
torch.cuda.synchronize()
for _ in range(2):
  _lib.subs_sample(lg, xs, M, 0.5, 0.49, U=U, out=cand)
  _lib.select_gather(sc, cand)
torch.cuda.synchronize()
print('done')
