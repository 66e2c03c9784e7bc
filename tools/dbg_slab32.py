"""Debug aid: pair-split conv + difference pooling, slab32 vs wide slabs, element-level diff."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import _lib  # noqa: E402

cuda = torch.device('cuda:0')
S, L, C = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (5, 100, 256)
g = torch.Generator().manual_seed(1)
A = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
res = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
W1 = (torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
bias = torch.randn(C, generator=g).to(cuda)
Wp = (2 * torch.eye(C) + torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
outs = {}
for mode in ('0', '1'):
  os.environ['SVDD_SLAB32'] = mode
  outs[mode] = [t.clone() for t in _lib.selftest_pair_pool(A, W1, bias, res, Wp, None, None, want_act=True)]
  torch.cuda.synchronize()
  print('mode', mode, 'ok')
for a, b, name in zip(outs['0'], outs['1'], ('y0', 'yd', 'pooled', 'pooled_act')):
  a2, b2 = a.float().reshape(-1, C), b.float().reshape(-1, C)
  bad = (a2 != b2) | torch.isnan(b2)
  print(name, 'equal' if not bool(bad.any()) else f'{int(bad.sum())} of {bad.numel()} differ')
  if bool(bad.any()):
    rows = bad.any(1).nonzero().flatten()[:12].tolist()
    cols = bad.any(0).nonzero().flatten()
    print('  rows', rows, 'cols', cols[:8].tolist(), '...', cols[-4:].tolist(), 'n_cols', int(cols.numel()))
    r = rows[0]
    print('  row', r, 'wide', a2[r, :8].tolist(), 'slab32', b2[r, :8].tolist())
    # is it a permutation of 8-column groups within the row?
    for u in range(4):
      for v in range(8):
        if torch.equal(a2[r, 8 * u:8 * u + 8], b2[r, 8 * v:8 * v + 8]):
          print(f'  wide unit {u} == slab32 unit {v}')
