"""Restatement of the five ``enformer_pytorch.modeling_enformer`` symbols that
the reference imports at ``Enformer.py:8-9`` -- TEST INFRASTRUCTURE ONLY.

Third-party dependency: ``enformer_pytorch`` (lucidrains/enformer-pytorch).
It is NOT vendored in /root/reference, NOT pinned anywhere in the reference
(absent from ``requirements.txt`` / ``requirements.yaml``; it arrives
transitively through ``gReLU``, itself un-pinned at ``requirements.txt:12``) and
NOT installed in this image.  What follows restates the published algorithm of
its ``modeling_enformer.py`` and is cross-checked against the reference's own
commented restatement of ``Attention`` at ``Enformer.py:2659-2768`` and against
the call sites:

  * ``Enformer.py:1843-1845``  exponential_linspace_int(768, 1536, num=6, divisible_by=128)
  * ``Enformer.py:1914-1923``  Attention(dim, heads, dim_key, dim_value, dropout,
                               pos_dropout, num_rel_pos_features, use_tf_gamma=False)
  * ``Enformer.py:2393``       GELU()
  * ``Enformer.py:2447``       AttentionPool(dim=in_channels, pool_size=pool_size)

PARITY UNPINNED at this boundary: the reference holds no test or golden vector
for these symbols and the real package is unavailable offline.  Parameter names
and shapes follow the published package so reference checkpoints would load.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F


def exponential_linspace_int(start, end, num, divisible_by=1):
  """Geometric progression start..end rounded to multiples of divisible_by."""
  ratio = math.exp(math.log(end / start) / (num - 1))
  out = []
  for i in range(num):
    out.append(int(round(start * ratio ** i / divisible_by) * divisible_by))
  return out


class GELU(nn.Module):
  """Enformer's sigmoid approximation: x * sigmoid(1.702 x)."""

  def forward(self, x):
    return torch.sigmoid(1.702 * x) * x


class AttentionPool(nn.Module):
  """Softmax-weighted pooling of ``pool_size`` adjacent positions per channel.

  Input [b, d, n] -> [b, d, ceil(n / pool_size)].  Logits come from a
  bias-free 1x1 Conv2d (initialised to 2*identity); a right-padded slot gets
  logit -finfo.max so it receives zero weight.
  """

  def __init__(self, dim, pool_size=2):
    super().__init__()
    self.pool_size = pool_size
    self.to_attn_logits = nn.Conv2d(dim, dim, 1, bias=False)
    nn.init.dirac_(self.to_attn_logits.weight)
    with torch.no_grad():
      self.to_attn_logits.weight.mul_(2)

  def forward(self, x):
    b, d, n = x.shape
    p = self.pool_size
    rem = n % p
    pad_mask = None
    if rem > 0:
      x = F.pad(x, (0, rem), value=0)
      pad_mask = torch.zeros((b, 1, n), dtype=torch.bool, device=x.device)
      pad_mask = F.pad(pad_mask, (0, rem), value=True)
      pad_mask = pad_mask.reshape(b, 1, -1, p)
    x = x.reshape(b, d, -1, p)
    logits = self.to_attn_logits(x)
    if pad_mask is not None:
      logits = logits.masked_fill(pad_mask, -torch.finfo(logits.dtype).max)
    w = logits.softmax(dim=-1)
    return (x * w).sum(dim=-1)


def _basis_exponential(dist, nfeat, seq_len):
  top = math.log(seq_len) / math.log(2.0)
  half_life = 2 ** torch.linspace(3.0, top, nfeat)
  return torch.exp(-math.log(2.0) / half_life[None, :] * dist.abs()[:, None])


def _basis_central_mask(dist, nfeat, seq_len):
  widths = 2 ** torch.arange(1, nfeat + 1).float() - 1
  return (widths[None, :] > dist.abs()[:, None]).float()


def _basis_gamma(dist, nfeat, seq_len, eps=1e-8):
  stddev = seq_len / (2 * nfeat)
  mean = torch.linspace(seq_len / nfeat, seq_len, nfeat)[None, :]
  conc = (mean / stddev) ** 2
  rate = mean / stddev ** 2
  x = dist.float().abs()[:, None]
  log_unnorm = torch.xlogy(conc - 1.0, x) - rate * x
  log_norm = torch.lgamma(conc) - conc * torch.log(rate)
  prob = torch.exp(log_unnorm - log_norm) + eps
  return prob / torch.amax(prob, dim=-1, keepdim=True)


def get_positional_embed(seq_len, feature_size, device=None, use_tf_gamma=False,
                         dtype=torch.float):
  """[2*seq_len-1, feature_size] relative-position basis (3 families, mirrored
  with sign)."""
  assert not use_tf_gamma, 'reference passes use_tf_gamma=False (Enformer.py:1922)'
  dist = torch.arange(-seq_len + 1, seq_len)
  if feature_size % 6 != 0:
    raise ValueError('feature size is not divisible by number of components (6)')
  per = feature_size // 6
  emb = torch.cat([_basis_exponential(dist, per, seq_len),
                   _basis_central_mask(dist, per, seq_len),
                   _basis_gamma(dist, per, seq_len)], dim=-1)
  emb = torch.cat([emb, torch.sign(dist)[:, None] * emb], dim=-1)
  emb = emb.to(dtype)
  return emb if device is None else emb.to(device)


def relative_shift(x):
  """Transformer-XL style shift of [.., t1, 2*t1-1] relative logits."""
  x = torch.cat([torch.zeros_like(x[..., :1]), x], dim=-1)
  _, h, t1, t2 = x.shape
  x = x.reshape(-1, h, t2, t1)[:, :, 1:, :]
  x = x.reshape(-1, h, t1, t2 - 1)
  return x[..., :((t2 + 1) // 2)]


class Attention(nn.Module):
  """Multi-head attention with Enformer relative positional logits.

  Follows the reference's commented restatement ``Enformer.py:2659-2768``:
  bias-free q/k/v, q scaled by dim_key^-0.5, content logits with
  ``rel_content_bias``, positional logits with ``rel_pos_bias`` over
  ``to_rel_k(positions)`` then ``relative_shift``, softmax, ``to_out``.
  """

  def __init__(self, dim, *, num_rel_pos_features, heads=8, dim_key=64,
               dim_value=64, dropout=0.0, pos_dropout=0.0, use_tf_gamma=False):
    super().__init__()
    self.scale = dim_key ** -0.5
    self.heads = heads
    self.to_q = nn.Linear(dim, dim_key * heads, bias=False)
    self.to_k = nn.Linear(dim, dim_key * heads, bias=False)
    self.to_v = nn.Linear(dim, dim_value * heads, bias=False)
    self.to_out = nn.Linear(dim_value * heads, dim)
    nn.init.zeros_(self.to_out.weight)
    nn.init.zeros_(self.to_out.bias)
    self.num_rel_pos_features = num_rel_pos_features
    self.to_rel_k = nn.Linear(num_rel_pos_features, dim_key * heads, bias=False)
    self.rel_content_bias = nn.Parameter(torch.randn(1, heads, 1, dim_key))
    self.rel_pos_bias = nn.Parameter(torch.randn(1, heads, 1, dim_key))
    self.pos_dropout = nn.Dropout(pos_dropout)
    self.attn_dropout = nn.Dropout(dropout)
    self.use_tf_gamma = use_tf_gamma

  def forward(self, x):
    b, n, _ = x.shape
    h = self.heads

    def split(t):
      return t.reshape(b, n, h, -1).permute(0, 2, 1, 3)

    q = split(self.to_q(x)) * self.scale
    k = split(self.to_k(x))
    v = split(self.to_v(x))
    content = torch.einsum('bhid,bhjd->bhij', q + self.rel_content_bias, k)
    pos = get_positional_embed(n, self.num_rel_pos_features, x.device,
                               use_tf_gamma=self.use_tf_gamma,
                               dtype=self.to_rel_k.weight.dtype)
    pos = self.pos_dropout(pos)
    rel_k = self.to_rel_k(pos).reshape(2 * n - 1, h, -1).permute(1, 0, 2)
    rel = torch.einsum('bhid,hjd->bhij', q + self.rel_pos_bias, rel_k)
    rel = relative_shift(rel)
    attn = self.attn_dropout((content + rel).softmax(dim=-1))
    out = torch.einsum('bhij,bhjd->bhid', attn, v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
    return self.to_out(out)
