"""Stress probe of the persistent tower kernel: many repetitions at several row counts, every
result compared bit for bit with the launch-per-GEMM path (debugging aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers
from svdd_b200 import value_nets
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for full in (True, False):
  emb, head = helpers.build_enformer(full=full)
  emb, head = emb.to(dev), head.to(dev)
  for n_cand in (128, 130, 700, 1280, 2048):
    tok = helpers.random_tokens(n_cand, 200, 7 + n_cand, 0.5).to(dev)
    os.environ['SVDD_TOWER'] = '0'
    ref = value_nets.score_tokens(emb, head, tok)
    os.environ['SVDD_TOWER'] = '1'
    os.environ['SVDD_TOWER_SPLITK'] = os.environ.get('STRESS_SPLITK', '1')   # '1': bit-exact against the per-launch path
    if os.environ['SVDD_TOWER_SPLITK'] != '1':
      ref = value_nets.score_tokens(emb, head, tok)                              # split-K: compare against its own first run
    nd = 0
    for i in range(reps):
      got = value_nets.score_tokens(emb, head, tok)
      if not torch.equal(got, ref):
        nd += 1
        print(f'  MISMATCH full={full} n_cand={n_cand} rep={i}: {int((got != ref).sum())} scores differ, '
              f'max {float((got - ref).abs().max()):.3e}')
    bad += nd
    print(f'full={full} n_cand={n_cand}: {reps - nd}/{reps} bit-exact')
print('TOTAL MISMATCHES', bad)
