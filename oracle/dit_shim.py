"""CPU restatement of the two flash-attn entry points the reference's DiT calls
(models/dit.py:107-110 and :262-263) -- TEST INFRASTRUCTURE ONLY.

flash-attn (pinned ``flash-attn==2.5.6`` in the reference's requirements.yaml:36) ships CUDA /
Triton kernels only, so neither function can run in the CPU-only build container.  Both have a
published closed form, restated here in plain torch and patched into the imported reference
module by tests/golden/ref_import.py when the DiT goldens are generated:

  * ``flash_attn.layers.rotary.apply_rotary_emb_qkv_(qkv, cos, sin)`` with the default
    ``interleaved=False``: GPT-NeoX style rotation of the first ``2 * cos.shape[-1]`` dims of q
    and k (v untouched): (x1, x2) = halves; out = (x1 cos - x2 sin, x1 sin + x2 cos).
  * ``flash_attn.flash_attn_interface.flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens,
    max_seqlen, dropout_p, causal=False)``: exact softmax(q k^T / sqrt(d)) v per sequence.

Parity at these two symbols is unpinned in the sense of DESIGN.md section 2 (third-party code
not in /root/reference); everything around them is the reference's own module code.
"""
import torch


def apply_rotary_emb_qkv_(qkv, cos, sin, cos_k=None, sin_k=None, interleaved=False, seqlen_offsets=0,
                          num_heads_q=None):
  """qkv [b, s, 3, h, d]; cos, sin [s, rd/2].  In place on q and k, returns qkv."""
  assert not interleaved and cos_k is None and sin_k is None
  rd = 2 * cos.shape[-1]
  c = cos[None, :, None, :].to(qkv.dtype)
  s = sin[None, :, None, :].to(qkv.dtype)
  for i in (0, 1):
    x = qkv[:, :, i]
    x1, x2 = x[..., :rd // 2].clone(), x[..., rd // 2:rd].clone()
    x[..., :rd // 2] = x1 * c - x2 * s
    x[..., rd // 2:rd] = x1 * s + x2 * c
  return qkv


def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None,
                                     causal=False, **kw):
  """qkv [total, 3, h, d] -> [total, h, d]; sequences delimited by cu_seqlens."""
  assert dropout_p == 0.0 and not causal
  d = qkv.shape[-1]
  scale = d ** -0.5 if softmax_scale is None else softmax_scale
  out = torch.empty_like(qkv[:, 0])
  cu = [int(v) for v in cu_seqlens]
  for a, b in zip(cu[:-1], cu[1:]):
    q, k, v = (qkv[a:b, i].transpose(0, 1).float() for i in range(3))     # [h, s, d]
    p = torch.softmax(torch.matmul(q, k.transpose(1, 2)) * scale, dim=-1)
    out[a:b] = torch.matmul(p, v).transpose(0, 1).to(out.dtype)
  return out
