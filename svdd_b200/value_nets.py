"""Value / reward networks of the SVDD decode path -- parameter containers whose
forward runs the sm_100a kernels.

Mirrors, for the configurations ``decode.py`` builds, the reference's
  * ``EnformerTrunk``  (Enformer.py:1271-1334)  DNA value net / reward-oracle family
  * ``ConvGRUTrunk``   (Enformer.py:1337-1426)  RNA value net / reward-oracle family
  * ``ConvHead``       (Enformer.py:2131-2173)
  * ``OriBaseModel``   (Enformer.py:1105-1127)  embedding+head wrapper used as reward model
The containers own the same parameter tree (state_dict keys and shapes) and are
built in the reference's construction order, so reference ``.pt`` checkpoints
load and ``torch.manual_seed`` reproduces the reference's random init.  They
hold no torch forward: calling them goes through the C ABI
(``svdd_value_convgru_*`` / ``svdd_value_enformer_*``); without the CUDA library
or on CPU tensors they raise.

``enformer_pytorch`` (third party, un-vendored) supplies ``Attention`` and
``AttentionPool`` parameters in the reference; their parameter layout is
restated here (to_q/to_k/to_v/to_out/to_rel_k/rel_content_bias/rel_pos_bias;
to_attn_logits initialised to 2*identity).
"""
import math

import torch
from torch import nn

from . import _lib


def _holder(**children):
  m = nn.Module()
  for name, child in children.items():
    m.add_module(name, child)
  return m


def _norm(kind, dim):
  """``Norm`` wrapper (Enformer.py:2507-2558): parameters live under ``.layer``."""
  layer = {'batch': nn.BatchNorm1d, 'layer': nn.LayerNorm}[kind](dim)
  return _holder(layer=layer)


def exponential_linspace_int(start, end, num, divisible_by=1):
  """enformer_pytorch helper used at Enformer.py:1843-1845."""
  base = math.exp(math.log(end / start) / (num - 1))
  return [int(round(start * base ** i / divisible_by) * divisible_by)
          for i in range(num)]


class ConvHead(nn.Module):
  """1x1 conv to n_tasks + average over length (Enformer.py:2131-2173) for the
  decode path's settings (norm=False, act_func=None, pool_func='avg')."""

  def __init__(self, n_tasks, in_channels, act_func=None, pool_func=None,
               norm=False):
    super().__init__()
    if act_func is not None or norm or pool_func != 'avg':
      raise NotImplementedError(
          "decode path uses ConvHead(act_func=None, norm=False, pool_func='avg')")
    self.n_tasks, self.in_channels = n_tasks, in_channels
    self.channel_transform = _holder(
        conv=_holder(layer=nn.Conv1d(in_channels, n_tasks, kernel_size=1)))

  def forward(self, x):
    """``head(embedding(onehot))`` as the reference spells it (diffusion_gosai.py:1208-1209,
    Enformer.py:443): ``x`` is what the trunk's ``forward`` returned -- a deferred handle on the
    token rows, because trunk and head run as ONE fused scoring pass.  Returns fp32
    [N, n_tasks, 1] like the reference's ConvHead (:2166-2173); with n_tasks > 1 every task is a
    scoring pass of its own (the engine's own path scores task 0 once, diffusion_gosai.py:1430)."""
    if not isinstance(x, _TrunkOutput):
      raise RuntimeError(
          'ConvHead is fused into the trunk kernels: it consumes the output of '
          'EnformerTrunk / ConvGRUTrunk.forward (or call score_tokens(embedding, head, tokens))')
    cols = [packed_scorer(x.trunk, self, task=t).score(x.tokens) for t in range(self.n_tasks)]
    return torch.stack(cols, dim=1).unsqueeze(-1)


class _TrunkOutput:
  """What ``EnformerTrunk.forward`` / ``ConvGRUTrunk.forward`` return: the trunk and the token rows
  recovered from the one-hot input.  The feature tensor itself never exists outside the kernels."""

  def __init__(self, trunk, tokens):
    self.trunk, self.tokens = trunk, tokens


def _onehot_to_tokens(x, channels_last):
  """fp one-hot [N,L,4] (or [N,4,L]) with all-zero rows for masked positions
  (transform_samples, diffusion_gosai.py:1462-1470) -> int64 tokens [N,L] with 4 = mask."""
  if not channels_last:
    x = x.transpose(1, 2)
  if x.dim() != 3 or x.shape[-1] != 4:
    raise ValueError(f'expected a one-hot input with 4 channels, got {tuple(x.shape)}')
  if not x.is_cuda:
    raise _lib.SvddError('svdd_b200 kernels need CUDA tensors (there is no CPU path)')
  hot = x != 0
  if bool(((x != 0) & (x != 1)).any()) or bool((hot.sum(-1) > 1).any()):
    raise ValueError('the scoring kernels consume token rows: the input must be one-hot '
                     '(soft inputs are not on the decode path)')
  return torch.where(hot.any(-1), x.argmax(-1), torch.full(x.shape[:2], 4, device=x.device))


class ConvGRUTrunk(nn.Module):
  """RNA value net trunk: Conv(4->C,k15)+ReLU -> (n_conv-1) x [Conv(C->C,k5) ->
  BN -> +res -> ReLU] -> biGRU(C) summed -> LN -> Linear(C,2C) -> ReLU ->
  Linear(2C,C).  Construction mirrors Enformer.py:1359-1409 (ConvTower :1634-1751,
  Stem :1754-1804, ConvBlock order 'CDNRA' :2176-2292, GRUBlock :1571-1630,
  FeedForwardBlock :2010-2047)."""

  def __init__(self, stem_in_channels=6, stem_channels=16, stem_kernel_size=15,
               n_conv=2, channel_init=16, channel_mult=1, kernel_size=5,
               act_func='relu', conv_norm=False, pool_func=None, pool_size=None,
               residual=False, crop_len=0, n_gru=1, dropout=0.0, gru_norm=False):
    super().__init__()
    if (act_func != 'relu' or pool_func is not None or crop_len != 0
        or n_gru != 1 or channel_mult != 1 or stem_channels != channel_init):
      raise NotImplementedError(
          'ConvGRUTrunk kernels cover the decode-path configuration '
          '(relu, no pooling/cropping, one GRU layer, constant width)')
    if conv_norm != residual:
      raise NotImplementedError('conv_norm and residual must be set together')
    C = stem_channels
    self.channels, self.n_conv = C, n_conv
    self.stem_kernel_size, self.kernel_size = stem_kernel_size, kernel_size
    self.conv_norm, self.residual = conv_norm, residual
    blocks = nn.ModuleList()
    # Stem: conv then an (unused) LayerNorm that still lives in the state_dict
    blocks.append(_holder(
        conv=nn.Conv1d(stem_in_channels, C, stem_kernel_size, padding='same'),
        norm=_norm('layer', C)))
    for _ in range(1, n_conv):
      if conv_norm:
        blocks.append(_holder(norm=_norm('batch', C),
                              conv=nn.Conv1d(C, C, kernel_size, padding='same')))
      else:
        blocks.append(_holder(conv=nn.Conv1d(C, C, kernel_size, padding='same')))
    self.conv_tower = _holder(blocks=blocks)
    gru = nn.GRU(input_size=C, hidden_size=C, bidirectional=True,
                 batch_first=True, num_layers=1)
    ffn = _holder(
        dense1=_holder(norm=_norm('layer', C), linear=nn.Linear(C, 2 * C)),
        dense2=_holder(linear=nn.Linear(2 * C, C)),
        dense=_holder(norm=_norm('layer', C), linear=nn.Linear(C, C)))  # unused
    self.gru_tower = _holder(gru=gru, ffn=ffn)

  def forward(self, x):
    """x: one-hot [N, L, 4] or [N, 4, L] (transposed when dim 1 is not the channel count,
    Enformer.py:1422-1423).  Returns the deferred trunk output ``ConvHead.forward`` consumes."""
    return _TrunkOutput(self, _onehot_to_tokens(x, channels_last=(x.shape[1] != 4)))


class _AttentionParams(nn.Module):
  """Parameter layout of enformer_pytorch ``Attention`` as constructed at
  Enformer.py:1914-1923 (RNG order: to_q, to_k, to_v, to_out, to_rel_k,
  rel_content_bias, rel_pos_bias; to_out zero-initialised)."""

  def __init__(self, dim, heads, dim_key, dim_value, num_rel_pos_features):
    super().__init__()
    self.heads, self.dim_key, self.dim_value = heads, dim_key, dim_value
    self.num_rel_pos_features = num_rel_pos_features
    self.to_q = nn.Linear(dim, dim_key * heads, bias=False)
    self.to_k = nn.Linear(dim, dim_key * heads, bias=False)
    self.to_v = nn.Linear(dim, dim_value * heads, bias=False)
    self.to_out = nn.Linear(dim_value * heads, dim)
    nn.init.zeros_(self.to_out.weight)
    nn.init.zeros_(self.to_out.bias)
    self.to_rel_k = nn.Linear(num_rel_pos_features, dim_key * heads, bias=False)
    self.rel_content_bias = nn.Parameter(torch.randn(1, heads, 1, dim_key))
    self.rel_pos_bias = nn.Parameter(torch.randn(1, heads, 1, dim_key))


def _attention_pool_params(dim):
  conv = nn.Conv2d(dim, dim, 1, bias=False)
  nn.init.dirac_(conv.weight)
  with torch.no_grad():
    conv.weight.mul_(2)
  return _holder(layer=_holder(to_attn_logits=conv))


def _nacdr(cin, cout, k, pool):
  """ConvBlock(order='NACDR') parameter holder: BN(cin) -> conv -> [attn pool]."""
  parts = dict(norm=_norm('batch', cin),
               conv=nn.Conv1d(cin, cout, k, padding='same'))
  if pool:
    parts['pool'] = _attention_pool_params(cout)
  return _holder(**parts)


class EnformerTrunk(nn.Module):
  """DNA value net trunk (Enformer.py:1271-1334): conv tower with attention
  pooling (EnformerConvTower :1807-1884), pre-LN transformer tower with
  Enformer relative-position attention (:1887-2007), pointwise ConvBlock and
  GELU.  Dropouts are identities in eval and hold no parameters."""

  def __init__(self, n_conv=7, channels=1536, n_transformers=11, n_heads=8,
               key_len=64, attn_dropout=0.05, pos_dropout=0.01, ff_dropout=0.4,
               crop_len=0):
    super().__init__()
    if crop_len != 0:
      raise NotImplementedError('crop_len must be 0 on the decode path')
    self.n_conv, self.channels = n_conv, channels
    self.n_transformers, self.n_heads, self.key_len = n_transformers, n_heads, key_len
    half = channels // 2
    self.filters = [half] + exponential_linspace_int(
        half, channels, num=n_conv - 1, divisible_by=128)
    blocks = nn.ModuleList()
    blocks.append(nn.Sequential(nn.Conv1d(4, half, 15, padding='same'),
                                _nacdr(half, half, 1, pool=True)))
    for i in range(1, n_conv):
      blocks.append(nn.Sequential(
          _nacdr(self.filters[i - 1], self.filters[i], 5, pool=False),
          _nacdr(self.filters[i], self.filters[i], 1, pool=True)))
    self.conv_tower = _holder(blocks=blocks)
    tblocks = nn.ModuleList()
    for _ in range(n_transformers):
      tblocks.append(_holder(
          norm=_norm('layer', channels),
          mha=_AttentionParams(channels, n_heads, key_len, channels // n_heads,
                               channels // n_heads),
          ffn=_holder(
              dense1=_holder(norm=_norm('layer', channels),
                             linear=nn.Linear(channels, 2 * channels)),
              dense2=_holder(linear=nn.Linear(2 * channels, channels)),
              dense=_holder(norm=_norm('layer', channels),
                            linear=nn.Linear(channels, channels)))))  # unused
    self.transformer_tower = _holder(blocks=tblocks)
    self.pointwise_conv = _nacdr(channels, 2 * channels, 1, pool=False)

  def forward(self, x):
    """x: one-hot [N, L, 4] (this trunk transposes unconditionally, Enformer.py:1328).  Returns the
    deferred trunk output ``ConvHead.forward`` consumes."""
    return _TrunkOutput(self, _onehot_to_tokens(x, channels_last=True))


class OriBaseModel(nn.Module):
  """embedding + head wrapper (Enformer.py:1105-1127); the reward-model shape
  ``reward_model(onehot.float().transpose(1, 2))[:, 0]`` of
  diffusion_gosai.py:1430 is served by ``score_tokens`` on token ids."""

  def __init__(self, embedding, head):
    super().__init__()
    self.embedding, self.head = embedding, head

  def forward(self, x):
    """x: one-hot [N, 4, L] -- the gReLU reward model's layout, which is why the path calls
    ``reward_model(onehot.float().transpose(1, 2))`` (diffusion_gosai.py:1430).  -> [N, n_tasks, 1]."""
    if isinstance(self.embedding, ConvGRUTrunk):
      return self.head(self.embedding(x))
    return self.head(_TrunkOutput(self.embedding, _onehot_to_tokens(x, channels_last=False)))


# -- packing + scoring ---------------------------------------------------------

_PACK_CACHE_ATTR = '_svdd_packed'


def packed_scorer(embedding, head, task=0):
  """Returns (and caches on ``embedding``) the packed device weights of an
  (embedding, head) pair for one task of the head.  Accepts the containers above or any
  nn.Module with the reference's state_dict layout (duck-typed on key names)."""
  sd = embedding.state_dict()
  key = tuple((v.data_ptr(), v._version) for v in sd.values()) + tuple(
      (v.data_ptr(), v._version) for v in head.state_dict().values())
  cache = getattr(embedding, _PACK_CACHE_ATTR, None)
  if cache is None or cache[0] != key:
    cache = (key, {})
    object.__setattr__(embedding, _PACK_CACHE_ATTR, cache)
  handle = cache[1].get(task)
  if handle is not None:
    return handle
  if any(k.startswith('gru_tower.') for k in sd):
    handle = _lib.ConvGRUHandle(sd, head.state_dict(), task=task)
  elif any(k.startswith('transformer_tower.') or k.startswith('conv_tower.blocks.0.0.')
           for k in sd):
    n_heads = getattr(embedding, 'n_heads', 8)
    handle = _lib.EnformerHandle(sd, head.state_dict(), n_heads, task=task)
  else:
    raise TypeError('unrecognised value-network parameter layout')
  cache[1][task] = handle
  return handle


def score_tokens(embedding, head, tokens):
  """head(embedding(transform_samples(tokens).float())).squeeze() of
  diffusion_gosai.py:1208-1209 for a batch of token rows.

  tokens: integer tensor [N, L] on a CUDA device, values 0..4 (4 = mask ->
  all-zero one-hot row, diffusion_gosai.py:1462-1470).  Returns fp32 [N]."""
  return packed_scorer(embedding, head).score(tokens)
