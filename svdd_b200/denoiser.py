"""MDLM denoiser backbone (the dilated CNN) -- parameter container + CUDA forward.

Mirror of the reference's ``models/dnaconv.py:135-210`` ``CNNModel`` for the
configuration the decode path uses (``clean_data=False``,
``cls_free_guidance=False``, ``classifier=False``, ``dropout=0``;
configs_gosai/model/dnaconv.yaml).  The module owns an identical parameter tree
(``linear``, ``time_embedder.{0.W,1}``, ``convs.i``, ``time_layers.i.dense``,
``norms.i``, ``final_conv.{0,2}``) built in the reference's construction order,
so (a) a reference Lightning checkpoint's ``backbone.*`` keys load with
``load_state_dict`` and (b) ``torch.manual_seed(s)`` yields the same random-init
weights as the reference constructor.

There is no torch forward: ``forward`` packs the weights once (bf16, tap-major,
time-bias folded) and runs the sm_100a kernels through the C ABI
(``svdd_denoiser_*`` in include/svdd_b200.h).  It raises if the CUDA library is
missing or the input is not on a CUDA device -- no CPU fallback.
"""
import copy
import math

import torch
from torch import nn

from . import _lib

DILATION_GROUPS = (1, 1, 4, 16, 64)   # models/dnaconv.py:155-160 (grouped, not cycled)


class GaussianFourierProjection(nn.Module):
  """Fixed random Fourier features of the time step (models/dnaconv.py:8-21)."""

  def __init__(self, embed_dim, scale=30.0):
    super().__init__()
    self.W = nn.Parameter(torch.randn(embed_dim // 2) * scale,
                          requires_grad=False)


class Dense(nn.Module):
  """``time_layers.i.dense`` holder (models/dnaconv.py:24-34)."""

  def __init__(self, input_dim, output_dim):
    super().__init__()
    self.dense = nn.Linear(input_dim, output_dim)


class CNNModel(nn.Module):
  """Denoiser: one-hot(5) -> Conv(5->H,k9)+ReLU -> n x [+time-bias -> LN(H) ->
  dilated Conv(H->H,k9) -> ReLU -> +residual] -> 1x1 -> ReLU -> 1x1 -> [B,L,5].
  """

  def __init__(self, args, alphabet_size, num_cls=3, classifier=False):
    super().__init__()
    if classifier or getattr(args, 'clean_data', False) or getattr(
        args, 'cls_free_guidance', False):
      raise NotImplementedError(
          'only the decode-path configuration of CNNModel is built '
          '(clean_data=False, cls_free_guidance=False, classifier=False)')
    if getattr(args, 'dropout', 0.0) != 0.0:
      raise NotImplementedError('dropout must be 0 on the decode path')
    self.args = args
    self.alphabet_size = alphabet_size
    self.num_cls = num_cls
    H = args.hidden_dim
    stacks = args.num_cnn_stacks
    self.num_layers = 5 * stacks
    self.linear = nn.Conv1d(alphabet_size, H, kernel_size=9, padding=4)
    self.time_embedder = nn.Sequential(GaussianFourierProjection(embed_dim=H),
                                       nn.Linear(H, H))
    protos = [nn.Conv1d(H, H, kernel_size=9, dilation=d, padding=4 * d)
              for d in DILATION_GROUPS]
    self.convs = nn.ModuleList(
        [copy.deepcopy(p) for p in protos for _ in range(stacks)])
    self.time_layers = nn.ModuleList(
        [Dense(H, H) for _ in range(self.num_layers)])
    self.norms = nn.ModuleList(
        [nn.LayerNorm(H) for _ in range(self.num_layers)])
    self.final_conv = nn.Sequential(nn.Conv1d(H, H, kernel_size=1), nn.ReLU(),
                                    nn.Conv1d(H, alphabet_size, kernel_size=1))
    self._packed = None
    self._packed_key = None

  # -- host-side weight preparation ------------------------------------------
  def dilations(self):
    stacks = self.num_layers // 5
    return [DILATION_GROUPS[i // stacks] for i in range(self.num_layers)]

  @torch.no_grad()
  def time_bias(self, sigma):
    """Per-layer time bias rows [num_layers, H] for a scalar sigma:
    time_layers[i](relu(time_embedder(sigma))) (models/dnaconv.py:182,192).
    sigma is a python float; all sequences of a batch share it on this path
    (0 when time_conditioning is False: diffusion_gosai.py:334-335)."""
    W = self.time_embedder[0].W.float()
    proj = torch.tensor(float(sigma), dtype=torch.float32,
                        device=W.device) * W * 2 * math.pi
    four = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
    emb = torch.relu(self.time_embedder[1](four[None]).float())
    rows = [tl.dense(emb)[0] for tl in self.time_layers]
    return torch.stack(rows, dim=0).contiguous()

  def _param_version(self):
    return tuple((p.data_ptr(), p._version) for p in self.parameters())

  def packed(self):
    """Device-side packed weights (svdd_denoiser_pack), rebuilt when any
    parameter storage changes (e.g. after load_state_dict / .cuda())."""
    key = self._param_version()
    if self._packed is None or self._packed_key != key:
      self._packed = _lib.DenoiserHandle(self)
      self._packed_key = key
    return self._packed

  # -- forward ---------------------------------------------------------------
  def forward(self, seq, t, cls=None, return_embedding=False):
    """seq int64[B,L] (CUDA) , t fp32[B] -> logits fp32[B,L,alphabet].
    All entries of ``t`` must be equal (true on the decode path)."""
    if cls is not None or return_embedding:
      raise NotImplementedError('cls / return_embedding are not on the decode path')
    sigma = float(t.reshape(-1)[0]) if t is not None and t.numel() else 0.0
    return self.packed().forward(seq, sigma)
