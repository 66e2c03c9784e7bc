"""Determinism / agreement probe of the persistent tower kernel (debugging aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers
from svdd_b200 import value_nets
dev = torch.device('cuda:0')
emb, head = helpers.build_enformer(full=True)
emb, head = emb.to(dev), head.to(dev)
for n_cand in (128, 1280):
  tok = helpers.random_tokens(n_cand, 200, 7 + n_cand, 0.5).to(dev)
  os.environ['SVDD_TOWER'] = '0'
  ref = value_nets.score_tokens(emb, head, tok).cpu()
  ref2 = value_nets.score_tokens(emb, head, tok).cpu()
  print(n_cand, 'per-launch path deterministic:', bool(torch.equal(ref, ref2)))
  os.environ['SVDD_TOWER'] = '1'
  for fast in ('0', '1'):
    os.environ['SVDD_TOWER_ATTN_FAST'] = fast
    outs = [value_nets.score_tokens(emb, head, tok).cpu() for _ in range(6)]
    nd = [int((o != outs[0]).sum()) for o in outs]
    err = [float((o - ref).abs().max()) for o in outs]
    print(n_cand, 'attn_fast', fast, 'differs-from-run0 counts', nd, 'max|d| vs per-launch', ['%.2e' % e for e in err])
