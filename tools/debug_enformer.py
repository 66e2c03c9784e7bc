"""Debug aid (GPU box): dumps the Enformer kernels' intermediates and compares them
stage by stage with the oracle's bf16-emulating forward.
    SVDD_DEBUG_DUMP_DIR=/tmp/efdump python tools/debug_enformer.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers
from oracle import nets, svdd
from svdd_b200 import value_nets

d = os.environ.setdefault('SVDD_DEBUG_DUMP_DIR', '/tmp/efdump')
os.makedirs(d, exist_ok=True)
dev = torch.device('cuda:0')
g = helpers.load_golden('value_nets.npz')
tok = torch.from_numpy(g['enformer_tokens'])
emb, head = helpers.build_enformer()
rec, rec32 = {}, {}
with torch.no_grad():
  nets.enformer_trunk(emb.state_dict(), svdd.transform_samples(tok).float(), 8, True, rec)
  nets.enformer_trunk(emb.state_dict(), svdd.transform_samples(tok).float(), 8, False, rec32)
got = value_nets.score_tokens(emb.to(dev), head.to(dev), tok.to(dev)).cpu()
print('scores', got.numpy())


def load(name, shape, dtype):
  raw = np.fromfile(os.path.join(d, name + '.bin'), dtype=np.uint16 if dtype == 'bf16' else np.float32)
  t = torch.from_numpy(raw.astype(np.int32) << 16).view(torch.float32) if dtype == 'bf16' else torch.from_numpy(raw)
  return t.reshape(shape)


def cmp(name, got, ref, key=None):
  err = (got - ref).abs()
  line = (f'{name:8s} |ref|max={float(ref.abs().max()):.2f} got-emu mean={float(err.mean()):.5f} max={float(err.max()):.4f}')
  if key is not None:
    r32 = rec32[key]
    r32 = r32.permute(0, 2, 1) if r32.shape != ref.shape else r32
    line += (f' | got-fp32 mean={float((got - r32).abs().mean()):.5f} max={float((got - r32).abs().max()):.4f}'
             f' | emu-fp32 mean={float((ref - r32).abs().mean()):.5f} max={float((ref - r32).abs().max()):.4f}')
  print(line)


N = tok.shape[0]
filters = emb.filters
L = 200
cmp('x0', load('ef_x0', (N, L, filters[0]), 'bf16'), rec['x0'].permute(0, 2, 1), 'x0')
for i in range(7):
  C = filters[i]
  if i > 0:
    cmp(f'z{i}', load(f'ef_z{i}', (N, L, C), 'bf16'), rec[f'z{i}'].permute(0, 2, 1), f'z{i}')
  if os.path.exists(os.path.join(d, f'ef_y{i}.bin')):
    cmp(f'y{i}', load(f'ef_y{i}', (N, L, C), 'bf16'), rec[f'y{i}'].permute(0, 2, 1), f'y{i}')
  else:    # pair path: y0 = y[2j], yd = y[2j+1] - y[2j] at half length
    Lo = (L + 1) // 2
    y0, yd = load(f'ef_y0_{i}', (N, Lo, C), 'bf16'), load(f'ef_yd_{i}', (N, Lo, C), 'bf16')
    y = torch.stack([y0, y0 + yd], 2).reshape(N, 2 * Lo, C)[:, :L]
    cmp(f'y{i}', y, rec[f'y{i}'].permute(0, 2, 1), f'y{i}')
  L = (L + 1) // 2
  if i == 6:
    cmp('xt_in', load('ef_xt_in', (N, L, C), 'f32'), rec['y6_pooled'].permute(0, 2, 1), 'y6_pooled')
for j in range(2):
  cmp(f'xattn{j}', load(f'ef_xattn{j}', (N, L, filters[-1]), 'f32'), rec[f'xattn{j}'], f'xattn{j}')
  cmp(f'xt{j}', load(f'ef_xt{j}', (N, L, filters[-1]), 'f32'), rec[f'xt{j}'], f'xt{j}')
