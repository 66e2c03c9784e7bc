"""Shared test utilities (seeded model builders used by both the golden
generator and the tests, so weights never have to be stored in fixtures)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, GOLDEN):
  if p not in sys.path:
    sys.path.insert(0, p)

CONVGRU_VALUE_KW = dict(
    stem_in_channels=4, stem_channels=64, stem_kernel_size=15, n_conv=6,
    channel_init=64, channel_mult=1, kernel_size=5, act_func='relu',
    conv_norm=True, pool_func=None, pool_size=None, residual=True, crop_len=0,
    n_gru=1, dropout=0.1, gru_norm=True)            # Enformer.py:32-48
CONVGRU_ORACLE_KW = dict(stem_in_channels=4, stem_channels=64, n_conv=6,
                         channel_init=64)           # rna_MRL_oracle.py:38-44 (gReLU defaults)
CONVGRU_HEAD_KW = dict(n_tasks=1, in_channels=64, act_func=None,
                       pool_func='avg', norm=False)  # Enformer.py:49
ENFORMER_SMALL_KW = dict(n_conv=7, channels=384, n_transformers=2, n_heads=8,
                         key_len=64, attn_dropout=0.05, pos_dropout=0.01,
                         ff_dropout=0.4, crop_len=0)
ENFORMER_FULL_KW = dict(n_conv=7, channels=1536, n_transformers=11, n_heads=8,
                        key_len=64, attn_dropout=0.05, pos_dropout=0.01,
                        ff_dropout=0.4, crop_len=0)  # decode.py:78-79


def model_cfg(length=200, hidden_dim=128, num_cnn_stacks=4):
  return types.SimpleNamespace(name='dnaconv', type='cnn', length=length,
                               hidden_dim=hidden_dim,
                               num_cnn_stacks=num_cnn_stacks, dropout=0.0,
                               clean_data=False, cls_free_guidance=False)


@torch.no_grad()
def perturb_(module, seed):
  """Deterministically de-trivialises a freshly constructed network: BatchNorm
  running stats / affine terms away from (0,1,1,0), LayerNorm affine away from
  (1,0), and zero-initialised projections (enformer Attention.to_out) to small
  noise -- otherwise the BN / attention paths are vacuous under random init
  (SURVEY.md section 8(d)).  Walks the state_dict in sorted-key order so the
  reference module and the svdd_b200 container receive identical values."""
  g = torch.Generator().manual_seed(seed)
  sd = module.state_dict()
  for k in sorted(sd):
    v = sd[k]
    if not v.dtype.is_floating_point:
      continue
    if k.endswith('running_mean'):
      v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    elif k.endswith('running_var'):
      v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    elif 'norm.layer.weight' in k or (k.startswith('norms.') and k.endswith('weight')):
      v.copy_(1.0 + 0.2 * (torch.rand(v.shape, generator=g) - 0.5))
    elif 'norm.layer.bias' in k or (k.startswith('norms.') and k.endswith('bias')):
      v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    elif 'to_out.weight' in k:
      v.copy_(torch.randn(v.shape, generator=g) * (v.shape[1] ** -0.5))
    elif 'to_out.bias' in k:
      v.copy_(torch.randn(v.shape, generator=g) * 0.05)
  return module


def random_tokens(B, L, seed, p_mask=0.4):
  g = torch.Generator().manual_seed(seed)
  x = torch.randint(0, 4, (B, L), generator=g)
  m = torch.rand(B, L, generator=g) < p_mask
  return torch.where(m, torch.full_like(x, 4), x)


def build_denoiser(seed=44, length=200):
  from svdd_b200 import denoiser
  torch.manual_seed(seed)
  return denoiser.CNNModel(model_cfg(length), alphabet_size=5, num_cls=3)


def build_convgru_value(seed=3, perturb=7):
  from svdd_b200 import value_nets
  torch.manual_seed(seed)
  emb = value_nets.ConvGRUTrunk(**CONVGRU_VALUE_KW)
  head = value_nets.ConvHead(**CONVGRU_HEAD_KW)
  perturb_(emb, perturb)
  return emb.eval(), head.eval()


def build_convgru_oracle(seed=4):
  from svdd_b200 import value_nets
  torch.manual_seed(seed)
  emb = value_nets.ConvGRUTrunk(**CONVGRU_ORACLE_KW)
  head = value_nets.ConvHead(**CONVGRU_HEAD_KW)
  return emb.eval(), head.eval()


def build_enformer(seed=5, perturb=9, full=False):
  from svdd_b200 import value_nets
  kw = ENFORMER_FULL_KW if full else ENFORMER_SMALL_KW
  torch.manual_seed(seed)
  emb = value_nets.EnformerTrunk(**kw)
  head = value_nets.ConvHead(n_tasks=1, in_channels=2 * kw['channels'],
                             act_func=None, pool_func='avg')
  perturb_(emb, perturb)
  if not full:
    # calibrated BatchNorm statistics (see tests/golden/make_golden.py:ref_enformer_small)
    stats = load_golden('enformer_small_bn.npz')
    sd = emb.state_dict()
    with torch.no_grad():
      for k in stats.files:
        sd[k].copy_(torch.from_numpy(stats[k]))
  return emb.eval(), head.eval()


def load_golden(name):
  return np.load(os.path.join(GOLDEN, name))


def state_checksum(sd):
  """Order-independent fingerprint of a state_dict (float64 sums)."""
  tot, tot_abs = 0.0, 0.0
  for k in sorted(sd):
    if sd[k].dtype.is_floating_point:
      v = sd[k].double()
      tot += float(v.sum())
      tot_abs += float(v.abs().sum())
  return np.array([tot, tot_abs])
