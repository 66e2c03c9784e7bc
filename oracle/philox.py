"""numpy restatement of the counter-based uniform stream used by the CUDA
kernels when no noise tensor is injected -- TEST INFRASTRUCTURE ONLY.

Not reference behaviour: the reference draws ``torch.rand_like`` from torch's
global generator (diffusion_gosai.py:30-34).  The kernels need a stream that
(a) does not read B*L*5*M floats per step from HBM and (b) is independent of how
the batch is sharded across GPUs, so they use Philox4x32-10 (Salmon et al.,
SC'11; the same generator family torch/cuRAND use) keyed by the run seed with
a counter made of (position, GLOBAL row, candidate, step).  The statistical
contract is the reference's: i.i.d. U[0,1) with 24 random bits, one uniform per
(candidate, row, position, vocab entry).

Counter layout (must match svdd_b200/csrc/philox.cuh):
  draws     : ctr = (l, row, m | part << 16, step)          part 0 -> v = 0..3
                                                            part 1 -> v = 4 (word 0)
  selection : ctr = (m // 4, row, 0, step | 1 << 24)        word m % 4
  key       = (seed & 0xffffffff, seed >> 32)
  uniform   = (word >> 8) * 2**-24
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Vectorised Philox4x32 with 10 rounds.  All inputs broadcastable uint32
  arrays; returns four uint32 arrays."""
  c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
  c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
  k0 = np.uint32(k0)
  k1 = np.uint32(k1)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = _M0 * c0.astype(np.uint64)
      p1 = _M1 * c2.astype(np.uint64)
      hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
      lo0 = (p0 & _MASK).astype(np.uint32)
      hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
      lo1 = (p1 & _MASK).astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32(k0 + _W0)
      k1 = np.uint32(k1 + _W1)
  return c0, c1, c2, c3


def _to_uniform(words):
  return ((words >> np.uint32(8)).astype(np.float32)
          * np.float32(2.0 ** -24))


def _key(seed):
  seed = int(seed) & 0xFFFFFFFFFFFFFFFF
  return seed & 0xFFFFFFFF, seed >> 32


def draw_uniforms(seed, step, M, B, L, row_offset=0):
  """fp32 [M,B,L,5] uniforms for the candidate draws of one reverse step."""
  k0, k1 = _key(seed)
  m = np.arange(M, dtype=np.uint32)[:, None, None]
  b = (np.arange(B, dtype=np.uint32) + np.uint32(row_offset))[None, :, None]
  l = np.arange(L, dtype=np.uint32)[None, None, :]
  a = philox4x32_10(l, b, m, np.uint32(step), k0, k1)
  e = philox4x32_10(l, b, m | np.uint32(1 << 16), np.uint32(step), k0, k1)
  out = np.empty((M, B, L, 5), dtype=np.float32)
  for v in range(4):
    out[..., v] = _to_uniform(a[v])
  out[..., 4] = _to_uniform(e[0])
  return out


def select_uniforms(seed, step, B, M, row_offset=0):
  """fp32 [B,M] uniforms for the alpha > 0 selection draw of one step."""
  k0, k1 = _key(seed)
  nblk = (M + 3) // 4
  blk = np.arange(nblk, dtype=np.uint32)[None, :]
  b = (np.arange(B, dtype=np.uint32) + np.uint32(row_offset))[:, None]
  w = philox4x32_10(blk, b, np.uint32(0), np.uint32(step) | np.uint32(1 << 24),
                    k0, k1)
  out = np.stack([_to_uniform(x) for x in w], axis=-1).reshape(B, nblk * 4)
  return np.ascontiguousarray(out[:, :M])
