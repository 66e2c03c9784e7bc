"""CPU restatement of the networks on the SVDD decode path -- TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional forwards over a ``state_dict`` whose keys/shapes are the reference's
(so reference checkpoints, the reference modules built in
tests/golden/make_golden.py and the product's parameter containers in
``svdd_b200`` all interoperate).  torch CPU fp32, same operator order as the
reference so outputs match it bit-for-bit on CPU (checked against
tests/golden/*.npz).

``emulate_bf16=True`` rounds to bfloat16 at the points where the sm_100a kernels
do -- GEMM/conv operands (activations and weights) and the activations that
cross a kernel boundary in bf16 (block outputs, residual inputs, pooled
values) -- keeping fp32 accumulation, normalisation, attention and the GRU
recurrence; it is the tight-tolerance comparison target for the tensor-core
path (the fp32 result is the loose, documented-tolerance target).
"""
import math

import torch
import torch.nn.functional as F

from . import enformer_shim


def _q(t, on):
  """bf16 operand rounding (identity when emulation is off)."""
  return t.to(torch.bfloat16).to(torch.float32) if on else t


# ----------------------------------------------------------------------------
# Denoiser: models/dnaconv.py:135-210 (CNNModel)
# ----------------------------------------------------------------------------

DENOISER_DILATION_GROUPS = (1, 1, 4, 16, 64)  # models/dnaconv.py:155-160


def denoiser_time_embedding(sd, sigma, prefix='backbone.'):
  """relu(Linear(GaussianFourierProjection(t))) -- models/dnaconv.py:19-21,182."""
  w = sd[prefix + 'time_embedder.0.W']
  proj = sigma[:, None] * w[None, :] * 2 * math.pi
  four = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
  return F.relu(F.linear(four, sd[prefix + 'time_embedder.1.weight'],
                         sd[prefix + 'time_embedder.1.bias']))


def denoiser_logits(sd, tokens, sigma=None, prefix='backbone.',
                    emulate_bf16=False):
  """CNNModel.forward (models/dnaconv.py:176-210), clean_data=False,
  cls_free_guidance=False, classifier=False, dropout=0.

  tokens int64[B,L] in 0..4 -> logits fp32[B,L,5].  ``sigma`` fp32[B] (zeros
  when time_conditioning is False: diffusion_gosai.py:334-335).
  """
  B, L = tokens.shape
  if sigma is None:
    sigma = torch.zeros(B, dtype=torch.float32)
  n_layers = sum(1 for k in sd if k.startswith(prefix + 'convs.')
                 and k.endswith('.weight'))
  stacks = n_layers // 5
  alphabet = sd[prefix + 'linear.weight'].shape[1]
  hidden = sd[prefix + 'linear.weight'].shape[0]
  temb = denoiser_time_embedding(sd, sigma, prefix)
  x = F.one_hot(tokens, num_classes=alphabet).float().permute(0, 2, 1)
  feat = F.relu(F.conv1d(x, sd[prefix + 'linear.weight'],
                         sd[prefix + 'linear.bias'], padding=4))
  for i in range(n_layers):
    d = DENOISER_DILATION_GROUPS[i // stacks]
    h = feat.clone()
    h = h + F.linear(temb, sd[prefix + f'time_layers.{i}.dense.weight'],
                     sd[prefix + f'time_layers.{i}.dense.bias'])[:, :, None]
    h = F.layer_norm(h.permute(0, 2, 1), (hidden,),
                     sd[prefix + f'norms.{i}.weight'],
                     sd[prefix + f'norms.{i}.bias']).permute(0, 2, 1)
    h = F.relu(F.conv1d(_q(h, emulate_bf16),
                        _q(sd[prefix + f'convs.{i}.weight'], emulate_bf16),
                        sd[prefix + f'convs.{i}.bias'],
                        dilation=d, padding=4 * d))
    feat = h + feat
  y = F.relu(F.conv1d(_q(feat, emulate_bf16),
                      _q(sd[prefix + 'final_conv.0.weight'], emulate_bf16),
                      sd[prefix + 'final_conv.0.bias']))
  y = F.conv1d(y, sd[prefix + 'final_conv.2.weight'],
               sd[prefix + 'final_conv.2.bias'])
  return y.permute(0, 2, 1)


# ----------------------------------------------------------------------------
# Shared pieces of Enformer.py
# ----------------------------------------------------------------------------

def _bn_eval(x, sd, p, eps=1e-5):
  """nn.BatchNorm1d in eval mode on [N,C,L] (Enformer.py:2523-2558)."""
  return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'],
                      sd[p + 'weight'], sd[p + 'bias'], False, 0.0, eps)


def _ln_channels(x, sd, p):
  """Norm('layer') on [N,L,C] (Enformer.py:2546-2550 reduces over C)."""
  c = sd[p + 'weight'].shape[0]
  return F.layer_norm(x, (c,), sd[p + 'weight'], sd[p + 'bias'])


def conv_head(sd, x, emulate_bf16=False):
  """ConvHead.forward (Enformer.py:2166-2173) with norm=False, act_func=None,
  pool_func='avg': 1x1 conv to n_tasks then mean over length.  [N,C,L]->[N,T,1]
  """
  y = F.conv1d(x, sd['channel_transform.conv.layer.weight'],
               sd['channel_transform.conv.layer.bias'])
  return y.mean(dim=2, keepdim=True)


# ----------------------------------------------------------------------------
# RNA value net / reward oracle: ConvGRUTrunk (Enformer.py:1337-1426)
# ----------------------------------------------------------------------------

def convgru_trunk(sd, x, emulate_bf16=False, residual=None):
  """ConvGRUTrunk.forward (Enformer.py:1411-1426) for the BaseModel
  configuration (Enformer.py:32-48): stem Conv(4->C,k15)+ReLU
  (Stem.forward :1790-1804, norm skipped), then n_conv-1 ConvBlocks in order
  CDNRA (:2266-2292): conv k5 -> BN(eval) -> +residual -> ReLU; biGRU summed
  over directions (:1607-1630); FeedForwardBlock = LN -> Linear -> ReLU ->
  Linear (:2044-2047, :2092-2099).  Blocks without BN keys (gReLU oracle
  variant: conv_norm=False) are handled too; ``residual`` defaults to "same as
  conv_norm", which is what both configurations on the path use.

  x fp32 [N,L,4] or [N,4,L] -> [N,C,L].
  """
  e = emulate_bf16
  cin = sd['conv_tower.blocks.0.conv.weight'].shape[1]
  if x.shape[1] != cin:                       # Enformer.py:1422-1423
    x = x.transpose(1, 2)
  w = sd['conv_tower.blocks.0.conv.weight']
  x = _q(F.relu(F.conv1d(x, w, sd['conv_tower.blocks.0.conv.bias'],
                         padding=w.shape[2] // 2)), e)
  i = 1
  while f'conv_tower.blocks.{i}.conv.weight' in sd:
    p = f'conv_tower.blocks.{i}.'
    w = sd[p + 'conv.weight']
    y = F.conv1d(_q(x, e), _q(w, e), sd[p + 'conv.bias'],
                 padding=w.shape[2] // 2)
    has_bn = (p + 'norm.layer.running_mean') in sd
    if has_bn:
      y = _bn_eval(y, sd, p + 'norm.layer.')
    if has_bn if residual is None else residual:
      y = y + x
    x = _q(F.relu(y), e)
    i += 1
  # GRUBlock.forward (Enformer.py:1607-1630)
  seq = x.permute(0, 2, 1)                     # [N,L,C]
  hidden = sd['gru_tower.gru.weight_hh_l0'].shape[1]
  out = _bigru(sd, seq, hidden, e)
  y = out[:, :, :hidden] + out[:, :, hidden:]
  y = _ln_channels(y, sd, 'gru_tower.ffn.dense1.norm.layer.')
  y = F.relu(F.linear(_q(y, e), _q(sd['gru_tower.ffn.dense1.linear.weight'], e),
                      sd['gru_tower.ffn.dense1.linear.bias']))
  y = F.linear(_q(y, e), _q(sd['gru_tower.ffn.dense2.linear.weight'], e),
               sd['gru_tower.ffn.dense2.linear.bias'])
  return y.permute(0, 2, 1)


def _bigru(sd, seq, hidden, emulate_bf16):
  """Single-layer bidirectional nn.GRU, batch_first (torch gate order r,z,n).

  Uses torch's own GRU kernel when not emulating so the CPU result equals the
  reference's ``nn.GRU`` bit-for-bit; the emulated path spells the recurrence
  out (input projection operands rounded to bf16, recurrence in fp32 as in the
  CUDA kernel).
  """
  names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0',
           'weight_ih_l0_reverse', 'weight_hh_l0_reverse',
           'bias_ih_l0_reverse', 'bias_hh_l0_reverse']
  if not emulate_bf16:
    # the ATen kernel behind nn.GRU.forward, called functionally so that no
    # module construction consumes the global RNG
    flat = [sd['gru_tower.gru.' + n] for n in names]
    h0 = torch.zeros(2, seq.shape[0], hidden, dtype=seq.dtype)
    with torch.no_grad():
      return torch.gru(seq, h0, flat, True, 1, 0.0, False, True, True)[0]
  N, L, _ = seq.shape
  outs = []
  for suffix, order in (('', range(L)), ('_reverse', range(L - 1, -1, -1))):
    wih = sd['gru_tower.gru.weight_ih_l0' + suffix]
    whh = sd['gru_tower.gru.weight_hh_l0' + suffix]
    bih = sd['gru_tower.gru.bias_ih_l0' + suffix]
    bhh = sd['gru_tower.gru.bias_hh_l0' + suffix]
    gi_all = F.linear(_q(seq, True), _q(wih, True), bih)       # [N,L,3H]
    h = torch.zeros(N, hidden)
    out = torch.zeros(N, L, hidden)
    for t in order:
      gi = gi_all[:, t]
      gh = F.linear(h, whh, bhh)
      r = torch.sigmoid(gi[:, :hidden] + gh[:, :hidden])
      z = torch.sigmoid(gi[:, hidden:2 * hidden] + gh[:, hidden:2 * hidden])
      n = torch.tanh(gi[:, 2 * hidden:] + r * gh[:, 2 * hidden:])
      h = (1 - z) * n + z * h
      out[:, t] = h
    outs.append(out)
  return torch.cat(outs, dim=2)


def convgru_value(sd_embedding, sd_head, onehot, emulate_bf16=False):
  """head(embedding(onehot)) as called at diffusion_gosai.py:1208-1209 /
  Enformer.py:443.  onehot fp32 [N,L,4] -> fp32 [N,1,1]."""
  return conv_head(sd_head, convgru_trunk(sd_embedding, onehot, emulate_bf16))


# ----------------------------------------------------------------------------
# DNA value net / reward oracle: EnformerTrunk (Enformer.py:1271-1334)
# ----------------------------------------------------------------------------

def _gelu_enformer(x):
  return torch.sigmoid(1.702 * x) * x          # enformer_pytorch GELU


def _attention_pool(x, w, emulate_bf16):
  """AttentionPool(pool_size=2) (enformer_pytorch; see enformer_shim)."""
  N, C, L = x.shape
  if L % 2:
    x = F.pad(x, (0, 1), value=0)
  pairs = x.reshape(N, C, -1, 2)
  logits = F.conv2d(_q(pairs, emulate_bf16), _q(w, emulate_bf16))
  if L % 2:
    mask = torch.zeros(1, 1, pairs.shape[2], 2, dtype=torch.bool)
    mask[:, :, -1, 1] = True
    logits = logits.masked_fill(mask, -torch.finfo(logits.dtype).max)
  return (pairs * logits.softmax(dim=-1)).sum(dim=-1)


def _nacdr_block(sd, p, x, residual, pool, e, rec=None, tag=''):
  """ConvBlock.forward with order 'NACDR' (Enformer.py:2266-2292):
  BN -> GELU -> conv -> (dropout) -> +input; then optional attention pool."""
  w = sd[p + 'conv.weight']
  y = _gelu_enformer(_bn_eval(x, sd, p + 'norm.layer.'))
  y = F.conv1d(_q(y, e), _q(w, e), sd[p + 'conv.bias'], padding=w.shape[2] // 2)
  if residual:
    y = y + x
  y = _q(y, e)
  if rec is not None:
    rec[tag] = y
  if pool:
    y = _attention_pool(y, sd[p + 'pool.layer.to_attn_logits.weight'], e)
    if rec is not None:
      rec[tag + '_pooled'] = y
  return y


def enformer_trunk(sd, x, n_heads=8, emulate_bf16=False, rec=None):
  """EnformerTrunk.forward (Enformer.py:1326-1334).

  x fp32 [N,L,4] -> [N,2*channels,L/2^n_conv].  Conv tower
  (EnformerConvTower :1807-1884), transformer tower (:1887-2007: pre-LN MHA
  with Enformer relative positions + residual, then FeedForwardBlock + residual),
  pointwise ConvBlock 'NACDR' k1 (:1315-1322), GELU (:1323).
  """
  e = emulate_bf16
  x = x.transpose(1, 2)                                   # :1328
  w = sd['conv_tower.blocks.0.0.weight']
  x = _q(F.conv1d(x, _q(w, e), sd['conv_tower.blocks.0.0.bias'],
                  padding=w.shape[2] // 2), e)
  if rec is not None:
    rec['x0'] = x
  x = _nacdr_block(sd, 'conv_tower.blocks.0.1.', x, True, True, e, rec, 'y0')
  i = 1
  while f'conv_tower.blocks.{i}.0.conv.weight' in sd:
    x = _nacdr_block(sd, f'conv_tower.blocks.{i}.0.', x, False, False, e, rec, f'z{i}')
    x = _nacdr_block(sd, f'conv_tower.blocks.{i}.1.', x, True, True, e, rec, f'y{i}')
    i += 1
  x = x.permute(0, 2, 1)                                  # [N,n,C]
  j = 0
  while f'transformer_tower.blocks.{j}.mha.to_q.weight' in sd:
    p = f'transformer_tower.blocks.{j}.'
    h = _ln_channels(x, sd, p + 'norm.layer.')
    x = x + _enformer_attention(sd, p + 'mha.', h, n_heads, e)
    if rec is not None:
      rec[f'xattn{j}'] = x
    h = _ln_channels(x, sd, p + 'ffn.dense1.norm.layer.')
    h = F.relu(F.linear(_q(h, e), _q(sd[p + 'ffn.dense1.linear.weight'], e),
                        sd[p + 'ffn.dense1.linear.bias']))
    h = F.linear(_q(h, e), _q(sd[p + 'ffn.dense2.linear.weight'], e),
                 sd[p + 'ffn.dense2.linear.bias'])
    x = x + h
    if rec is not None:
      rec[f'xt{j}'] = x
    j += 1
  x = x.permute(0, 2, 1)
  x = _nacdr_block(sd, 'pointwise_conv.', x, False, False, e)
  return _gelu_enformer(x)


def _enformer_attention(sd, p, x, heads, e):
  """enformer_pytorch Attention.forward (restated in enformer_shim.Attention;
  reference's commented copy: Enformer.py:2659-2768)."""
  b, n, _ = x.shape
  dk = sd[p + 'to_q.weight'].shape[0] // heads
  feats = sd[p + 'to_rel_k.weight'].shape[1]

  def split(t):
    return t.reshape(b, n, heads, -1).permute(0, 2, 1, 3)

  xq = _q(x, e)
  q = split(F.linear(xq, _q(sd[p + 'to_q.weight'], e))) * dk ** -0.5
  k = split(F.linear(xq, _q(sd[p + 'to_k.weight'], e)))
  v = split(F.linear(xq, _q(sd[p + 'to_v.weight'], e)))
  content = torch.einsum('bhid,bhjd->bhij', q + sd[p + 'rel_content_bias'], k)
  pos = enformer_shim.get_positional_embed(n, feats)
  rel_k = F.linear(pos, sd[p + 'to_rel_k.weight'])
  rel_k = rel_k.reshape(2 * n - 1, heads, dk).permute(1, 0, 2)
  rel = torch.einsum('bhid,hjd->bhij', q + sd[p + 'rel_pos_bias'], rel_k)
  rel = enformer_shim.relative_shift(rel)
  attn = (content + rel).softmax(dim=-1)
  out = torch.einsum('bhij,bhjd->bhid', attn, v)
  out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
  return F.linear(_q(out, e), _q(sd[p + 'to_out.weight'], e),
                  sd[p + 'to_out.bias'])


def enformer_value(sd_embedding, sd_head, onehot, n_heads=8,
                   emulate_bf16=False):
  """head(embedding(onehot)): Enformer.py:443, diffusion_gosai.py:1208-1209.
  onehot fp32 [N,L,4] -> fp32 [N,1,1]."""
  return conv_head(sd_head, enformer_trunk(sd_embedding, onehot, n_heads,
                                           emulate_bf16))


# ----------------------------------------------------------------------------
# DiT denoiser: models/dit.py:214-369
# ----------------------------------------------------------------------------

def _dit_ln(x, w):
  """LayerNorm without bias (models/dit.py:126-134): F.layer_norm(x.float(), [dim]) * weight."""
  return F.layer_norm(x.float(), (x.shape[-1],)) * w[None, None, :]


def dit_logits(sd, tokens, sigma=None, n_heads=12, prefix='backbone.', emulate_bf16=False):
  """DIT.forward body (models/dit.py:355-366; the reference's function lacks its ``return``):
  vocab_embed -> c = silu(sigma_map(sigma)) -> DDiTBlock x n (:239-288) -> DDitFinalLayer
  (:316-321).  tokens int64 [B,L] -> logits fp32 [B,L,V].  fp32 throughout (the reference's
  bf16 autocast region is a GPU-only construct); ``emulate_bf16`` rounds GEMM operands, q / k / v,
  the attention probabilities and the branch outputs to bf16 where the sm_100a kernels do."""
  e = emulate_bf16
  B, L = tokens.shape
  if sigma is None:
    sigma = torch.zeros(B, dtype=torch.float32)
  g = lambda k: sd[prefix + k]
  x = g('vocab_embed.embedding')[tokens]                                 # EmbeddingLayer :296-297
  half = g('sigma_map.mlp.0.weight').shape[1] // 2
  freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
  args = sigma[:, None].float() * freqs[None]
  t_freq = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)         # :160-181
  t_emb = F.linear(F.silu(F.linear(t_freq, g('sigma_map.mlp.0.weight'), g('sigma_map.mlp.0.bias'))),
                   g('sigma_map.mlp.2.weight'), g('sigma_map.mlp.2.bias'))
  c = F.silu(t_emb)                                                      # :357
  inv_freq = g('rotary_emb.inv_freq')
  fr = torch.arange(L, dtype=torch.float32)[:, None] * inv_freq[None, :]  # :88-90
  cos, sin = fr.cos()[None, :, None, :], fr.sin()[None, :, None, :]
  hd = x.shape[-1] // n_heads
  i = 0
  while prefix + f'blocks.{i}.attn_qkv.weight' in sd:
    p = f'blocks.{i}.'
    mod = F.linear(c, g(p + 'adaLN_modulation.weight'), g(p + 'adaLN_modulation.bias'))[:, None]
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=2)                     # :245-246
    h = _dit_ln(x, g(p + 'norm1.weight')) * (1 + sc1) + sh1              # :250
    qkv = _q(F.linear(_q(h, e), _q(g(p + 'attn_qkv.weight'), e)), e)     # :252
    qkv = qkv.reshape(B, L, 3, n_heads, hd)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]

    def rot(t):
      t1, t2 = t[..., :hd // 2], t[..., hd // 2:]
      return _q(torch.cat([t1 * cos - t2 * sin, t1 * sin + t2 * cos], dim=-1), e)
    q, k = rot(q), rot(k)                                                # :257-260
    att = torch.einsum('bqhd,bkhd->bhqk', q, k) * hd ** -0.5
    if e:       # the kernel rounds exp(s - max) to bf16 for the P V product and normalises by the sum of the ROUNDED values
      pu = _q(torch.exp(att - att.amax(-1, keepdim=True)), True)
      pr = pu / pu.sum(-1, keepdim=True)
    else:
      pr = torch.softmax(att, dim=-1)
    a = torch.einsum('bhqk,bkhd->bqhd', pr, v).reshape(B, L, -1)         # :262-265
    x = x + g1 * F.linear(_q(a, e), _q(g(p + 'attn_out.weight'), e))     # :267-271
    h = _dit_ln(x, g(p + 'norm2.weight')) * (1 + sc2) + sh2
    u = F.gelu(F.linear(_q(h, e), _q(g(p + 'mlp.0.weight'), e), g(p + 'mlp.0.bias')), approximate='tanh')
    x = x + g2 * F.linear(_q(u, e), _q(g(p + 'mlp.2.weight'), e), g(p + 'mlp.2.bias'))   # :274-277
    i += 1
  mod = F.linear(c, g('output_layer.adaLN_modulation.weight'), g('output_layer.adaLN_modulation.bias'))[:, None]
  sh, sc = mod.chunk(2, dim=2)
  h = _dit_ln(x, g('output_layer.norm_final.weight')) * (1 + sc) + sh
  return F.linear(h, g('output_layer.linear.weight'), g('output_layer.linear.bias'))
