"""DiT denoiser backbone -- parameter container + CUDA forward.

Mirror of the reference's ``models/dit.py`` (``DIT`` :324-369, ``DDiTBlock`` :214-288,
``DDitFinalLayer`` :300-321, ``TimestepEmbedder`` :148-187, ``EmbeddingLayer`` :291-297,
``Rotary`` :74-99, ``LayerNorm`` :126-134), selected by ``backbone: dit``
(diffusion_gosai.py:102-104, configs_gosai/model/small.yaml).  The module owns the same
parameter tree, built in the reference's construction order, so ``backbone.*`` keys of a
reference checkpoint load and ``torch.manual_seed`` reproduces the reference's random init.

Two facts about the reference, both kept visible here rather than papered over:
  * ``models/__init__.py`` comments ``dit`` out and ``DIT.forward`` ends without ``return x``
    (:355-366) -- the backbone cannot run in the reference as shipped.  This
    container returns the logits of ``output_layer`` (what upstream MDLM returns), and the
    goldens drive the reference's own sub-modules in the order of that forward body.
  * the six adaLN vectors depend only on ``c = silu(sigma_map(sigma))``; on the decode path
    sigma is the same for every sequence of a call (0 without time conditioning,
    diffusion_gosai.py:334-335), so ``modulation(sigma)`` evaluates them once on the host and
    the kernels receive per-channel constants (``svdd_dit_forward``'s ``mod``).

There is no torch forward: ``forward`` runs the sm_100a kernels through the C ABI
(``svdd_dit_*``) and raises without the CUDA library or on CPU tensors.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from . import _lib


class LayerNorm(nn.Module):
  """Weight-only LayerNorm (models/dit.py:126-134)."""

  def __init__(self, dim):
    super().__init__()
    self.weight = nn.Parameter(torch.ones([dim]))
    self.dim = dim


class TimestepEmbedder(nn.Module):
  """Sinusoidal features -> Linear -> SiLU -> Linear (models/dit.py:148-187)."""

  def __init__(self, hidden_size, frequency_embedding_size=256):
    super().__init__()
    self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                             nn.Linear(hidden_size, hidden_size, bias=True))
    self.frequency_embedding_size = frequency_embedding_size

  @staticmethod
  def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
                      ).to(device=t.device)
    args = t[:, None].float() * freqs[None]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
      embedding = torch.cat([embedding, torch.zeros_like(embedding[:, :1])], dim=-1)
    return embedding

  def forward(self, t):
    return self.mlp(self.timestep_embedding(t, self.frequency_embedding_size))


class Rotary(nn.Module):
  def __init__(self, dim, base=10_000):
    super().__init__()
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    self.register_buffer('inv_freq', inv_freq)


class EmbeddingLayer(nn.Module):
  def __init__(self, dim, vocab_dim):
    super().__init__()
    self.embedding = nn.Parameter(torch.empty((vocab_dim, dim)))
    torch.nn.init.kaiming_uniform_(self.embedding, a=math.sqrt(5))


class DDiTBlock(nn.Module):
  def __init__(self, dim, n_heads, cond_dim, mlp_ratio=4, dropout=0.1):
    super().__init__()
    self.n_heads = n_heads
    self.norm1 = LayerNorm(dim)
    self.attn_qkv = nn.Linear(dim, 3 * dim, bias=False)
    self.attn_out = nn.Linear(dim, dim, bias=False)
    self.norm2 = LayerNorm(dim)
    self.mlp = nn.Sequential(nn.Linear(dim, mlp_ratio * dim, bias=True), nn.GELU(approximate='tanh'),
                             nn.Linear(mlp_ratio * dim, dim, bias=True))
    self.adaLN_modulation = nn.Linear(cond_dim, 6 * dim, bias=True)
    self.adaLN_modulation.weight.data.zero_()
    self.adaLN_modulation.bias.data.zero_()


class DDitFinalLayer(nn.Module):
  def __init__(self, hidden_size, out_channels, cond_dim):
    super().__init__()
    self.norm_final = LayerNorm(hidden_size)
    self.linear = nn.Linear(hidden_size, out_channels)
    self.linear.weight.data.zero_()
    self.linear.bias.data.zero_()
    self.adaLN_modulation = nn.Linear(cond_dim, 2 * hidden_size, bias=True)
    self.adaLN_modulation.weight.data.zero_()
    self.adaLN_modulation.bias.data.zero_()


class DIT(nn.Module):
  """config.model: hidden_size, cond_dim, n_blocks, n_heads, dropout, scale_by_sigma, length."""

  def __init__(self, config, vocab_size):
    super().__init__()
    self.config = config
    self.vocab_size = vocab_size
    m = config.model
    if m.hidden_size != 64 * m.n_heads or m.hidden_size % 128:
      raise NotImplementedError('the DiT kernels need head dim 64 and hidden_size % 128 == 0 '
                                '(configs_gosai/model/small.yaml: 768 / 12)')
    self.n_heads = m.n_heads
    self.vocab_embed = EmbeddingLayer(m.hidden_size, vocab_size)
    self.sigma_map = TimestepEmbedder(m.cond_dim)
    self.rotary_emb = Rotary(m.hidden_size // m.n_heads)
    self.blocks = nn.ModuleList([DDiTBlock(m.hidden_size, m.n_heads, m.cond_dim, dropout=m.dropout)
                                 for _ in range(m.n_blocks)])
    self.output_layer = DDitFinalLayer(m.hidden_size, vocab_size, m.cond_dim)
    self.scale_by_sigma = m.scale_by_sigma
    self._packed = None
    self._packed_key = None

  @torch.no_grad()
  def modulation(self, sigma):
    """Folded adaLN vectors for a scalar sigma, fp32 [(8 * n_blocks + 2) * H] on the parameters'
    device (layout: csrc/dit.cu, svdd_dit_forward).  c = silu(sigma_map(sigma)) (models/dit.py:357);
    per block (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp) =
    adaLN_modulation(c).chunk(6) (:245-246); modulate(LN(x) * w, shift, scale) =
    LN(x) * [w (1 + scale)] + shift (:119-120, :250, :281)."""
    dev = self.vocab_embed.embedding.device
    t = torch.tensor([float(sigma)], dtype=torch.float32, device=dev)
    c = F.silu(self.sigma_map(t).float())
    out = []
    for blk in self.blocks:
      sh1, sc1, g1, sh2, sc2, g2 = blk.adaLN_modulation(c)[0].float().chunk(6)
      out += [blk.norm1.weight.float() * (1 + sc1), sh1, g1, torch.zeros_like(g1),
              blk.norm2.weight.float() * (1 + sc2), sh2, g2, g2 * blk.mlp[2].bias.float()]
    sh, sc = self.output_layer.adaLN_modulation(c)[0].float().chunk(2)
    out += [self.output_layer.norm_final.weight.float() * (1 + sc), sh]
    return torch.cat(out).contiguous()

  # the DenoiserHandle interface of denoiser.CNNModel
  def time_bias(self, sigma):
    return self.modulation(sigma)

  def _param_version(self):
    return tuple((p.data_ptr(), p._version) for p in self.parameters())

  def packed(self):
    key = self._param_version()
    if self._packed is None or self._packed_key != key:
      self._packed = _lib.DiTHandle(self)
      self._packed_key = key
    return self._packed

  def forward(self, indices, sigma):
    """indices int64[B,L] (CUDA), sigma fp32[B] (all equal on the decode path) -> logits fp32[B,L,V]."""
    s = float(sigma.reshape(-1)[0]) if sigma is not None and sigma.numel() else 0.0
    return self.packed().forward(indices, s)
