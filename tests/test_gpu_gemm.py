"""-m gpu: the tcgen05/TMA implicit-GEMM building block against a plain
CUDA-core kernel of the same contraction and against torch fp32 conv1d."""
import pytest
import torch
import torch.nn.functional as F

from svdd_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [
    # S, L, K, N, taps, dil
    (3, 200, 128, 128, 9, 1),      # denoiser layer, dilation 1
    (3, 200, 128, 128, 9, 4),
    (2, 200, 128, 128, 9, 16),
    (2, 200, 128, 128, 9, 64),     # most taps fall in the padding and are skipped
    (5, 50, 128, 128, 9, 64),      # RNA length: only the centre tap is in range
    (7, 50, 64, 64, 5, 1),         # ConvGRU conv block
    (9, 100, 768, 768, 5, 1),      # Enformer tower conv (L=100)
    (11, 25, 896, 1024, 5, 1),     # odd-length tiles, several sequences per tile
    (40, 7, 128, 256, 5, 1),
    (64, 4, 1280, 1536, 5, 1),
    (4, 200, 768, 768, 1, 1),      # 1x1 conv == plain GEMM (flattened rows)
    (300, 2, 1536, 2560, 1, 1),    # fused QKV projection
    (300, 2, 3072, 1536, 1, 1),
    (1, 1, 64, 64, 1, 1),
    (80, 50, 256, 896, 5, 1),      # N = 3*256 + 128: two column windows (256- and 128-wide tiles)
    (60, 100, 128, 1152, 1, 1),
]


@pytest.mark.parametrize('S,L,K,N,taps,dil', CASES)
def test_conv_gemm_matches_cuda_core_kernel(cuda, S, L, K, N, taps, dil):
  g = torch.Generator(device='cpu').manual_seed(S * 131 + L)
  A = (torch.randn(S, L, K, generator=g)).to(cuda).bfloat16()
  W = (torch.randn(taps, N, K, generator=g) * (K * taps) ** -0.5).to(cuda).bfloat16()
  bias = torch.randn(N, generator=g).to(cuda)
  ref = _lib.selftest_conv_gemm(A, W, bias, taps, dil, tensor_cores=False)
  out = _lib.selftest_conv_gemm(A, W, bias, taps, dil, tensor_cores=True)
  torch.cuda.synchronize()
  err = float((out - ref).abs().max())
  assert err < 2e-3, err
  # and against torch's own conv1d in fp32 on the same bf16-rounded operands
  y = F.conv1d(A.float().permute(0, 2, 1), W.float().permute(1, 2, 0), bias,
               dilation=dil, padding=(taps // 2) * dil)
  y = y.permute(0, 2, 1).reshape(S * L, N)
  assert float((out - y).abs().max()) < 5e-3


@pytest.mark.parametrize('S,L,C', [(3, 200, 128), (5, 100, 768), (7, 25, 256), (9, 13, 128),
                                   (20, 7, 384), (33, 4, 128), (1, 1, 64), (2, 3, 64)])
def test_attention_pool_epilogue(cuda, S, L, C):
  """EPI_POOL: two TMEM accumulators (even / odd positions), pair softmax and the
  weighted sum, incl. odd lengths where the last pair has a padded slot."""
  g = torch.Generator().manual_seed(S * 7 + L)
  y = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
  Wp = (2 * torch.eye(C) + torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
  out = _lib.selftest_pool(y, Wp)
  yf = y.float()
  logits = yf @ Wp.float().t()
  if L % 2:
    yf = F.pad(yf, (0, 0, 0, 1))
    logits = F.pad(logits, (0, 0, 0, 1), value=-torch.finfo(torch.float32).max)
  w = logits.reshape(S, -1, 2, C).softmax(dim=2)
  ref = (yf.reshape(S, -1, 2, C) * w).sum(2).reshape(-1, C)
  assert float((out - ref).abs().max()) < 2e-3


@pytest.mark.parametrize('S,L,C', [(3, 200, 128), (5, 100, 768), (7, 50, 256), (40, 25, 256), (9, 13, 128),
                                   (20, 7, 384), (33, 4, 128), (70, 1, 128), (2, 3, 256), (1300, 2, 128),
                                   (64, 50, 896), (200, 13, 1152)])
def test_pair_split_difference_pooling(cuda, S, L, C):
  """EPI_PAIR + EPI_POOL2 (the path svdd_enformer_score takes): the residual 1x1 conv emits
  y0 = y[2j] and yd = y[2j+1] - y[2j]; ONE GEMM over yd gives the pair softmax.  Checked
  against the reference formulation (two logits per pair, softmax, weighted sum; the padded
  slot of an odd length gets -max), incl. ragged tiles and odd lengths."""
  g = torch.Generator().manual_seed(S * 7 + L)
  A = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
  res = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
  W1 = (torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
  bias = torch.randn(C, generator=g).to(cuda)
  Wp = (2 * torch.eye(C) + torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
  scale2 = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda)
  shift2 = (0.3 * torch.randn(C, generator=g)).to(cuda)
  y0, yd, pooled, pact = _lib.selftest_pair_pool(A, W1, bias, res, Wp, scale2, shift2, want_act=True)
  torch.cuda.synchronize()
  y = A.float() @ W1.float().t() + bias + res.float()                     # [S, L, C] fp32
  Lo = (L + 1) // 2
  yp = F.pad(y, (0, 0, 0, 2 * Lo - L))
  ref_y0, ref_y1 = yp[:, 0::2], yp[:, 1::2]
  ref_yd = ref_y1 - ref_y0
  if L % 2:
    ref_yd[:, -1] = 0.0
  scale = max(1.0, float(y.abs().max()))
  assert float((y0.float() - ref_y0).abs().max()) < 2e-2 * scale
  assert float((yd.float() - ref_yd).abs().max()) < 2e-2 * scale
  # pooling semantics on the kernel's own (bf16) y0 / yd, as the reference formulates them
  y0f, y1f = y0.float(), y0.float() + yd.float()
  l0, l1 = y0f @ Wp.float().t(), y1f @ Wp.float().t()
  if L % 2:
    l1[:, -1] = -torch.finfo(torch.float32).max
  w = torch.stack([l0, l1], 2).softmax(dim=2)
  ref = (torch.stack([y0f, y1f], 2) * w).sum(2).reshape(-1, C)
  # the logit difference is formed from bf16(yd): |d err| <= 2^-9 |Wp.yd|, sigmoid slope <= 1/4
  assert float((pooled - ref).abs().max()) < 3e-2 * scale
  assert float((pooled - ref).abs().mean()) < 2e-3 * scale
  ref_act = _act(ref * scale2 + shift2, 2)
  assert float((pact.float() - ref_act).abs().max()) < 4e-2 * max(1.0, float(ref_act.abs().max()))


def test_relative_position_basis_matches_oracle():
  from oracle import enformer_shim
  for n, Fdim in [(2, 192), (2, 48), (4, 96), (7, 192)]:
    got = _lib.selftest_rel_positions(n, Fdim)
    ref = enformer_shim.get_positional_embed(n, Fdim)
    # fp32 torch vs fp64 host evaluation of the gamma log-pdf (lgamma of ~1e3)
    assert float((got - ref).abs().max()) < 2e-3, (n, Fdim)


@pytest.mark.parametrize('dv', [48, 192])
@pytest.mark.parametrize('n', [1, 2, 4])
def test_attention_kernel(cuda, n, dv, monkeypatch):
  """Both attention kernels (warp per (sequence, head); warp per sequence = the persistent
  tower's attention item) against the reference formulation, and against each other."""
  from oracle import enformer_shim
  H, dk, rows = 8, 64, 37
  g = torch.Generator().manual_seed(n)
  qkv = torch.randn(rows * n, 2 * H * dk + H * dv, generator=g)
  rcb, rpb = torch.randn(H * dk, generator=g), torch.randn(H * dk, generator=g)
  relk = torch.randn(H, 2 * n - 1, dk, generator=g) * 0.3
  q = qkv[:, :H * dk].reshape(rows, n, H, dk).permute(0, 2, 1, 3) * dk ** -0.5
  k = qkv[:, H * dk:2 * H * dk].reshape(rows, n, H, dk).permute(0, 2, 1, 3)
  v = qkv[:, 2 * H * dk:].reshape(rows, n, H, dv).permute(0, 2, 1, 3)
  content = torch.einsum('bhid,bhjd->bhij', q + rcb.reshape(1, H, 1, dk), k)
  rel = enformer_shim.relative_shift(torch.einsum('bhid,hjd->bhij', q + rpb.reshape(1, H, 1, dk), relk))
  out = torch.einsum('bhij,bhjd->bhid', (content + rel).softmax(-1), v)
  ref = out.permute(0, 2, 1, 3).reshape(rows * n, H * dv)
  got = {}
  for seq in ('0', '1'):
    monkeypatch.setenv('SVDD_ATTN_SEQ', seq)
    got[seq] = _lib.selftest_attention(qkv.to(cuda), rcb.to(cuda), rpb.to(cuda), relk.to(cuda), n, H, dk, dv).float().cpu()
    # bf16 output: half an ulp (2^-9 relative) + softmax with __expf
    err = float((got[seq] - ref).abs().max())
    print(f'\n[attention n={n} dv={dv} seq={seq}] max|d|={err:.3e} scale={float(ref.abs().max()):.3e}')
    assert err < 6e-3 * float(ref.abs().max())
  assert float((got['0'] - got['1']).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())   # <= 1 bf16 ulp apart


def _act(v, kind):
  if kind == 1:
    return v.relu()
  if kind == 2:
    return v * torch.sigmoid(1.702 * v)
  return v


EPI_CASES = [
    # S, L, K, N, taps, flat, out dtype, res?, dual?, act, act_after_res, scale?
    (9, 200, 64, 768, 1, True, torch.bfloat16, False, True, 0, False, False),     # stem: x0 + GELU(BN(x0))
    (5, 100, 256, 768, 5, False, torch.bfloat16, False, True, 0, False, False),   # k5 conv, two outputs
    (7, 200, 768, 768, 1, True, torch.bfloat16, True, False, 0, False, False),    # 1x1 + bf16 residual
    (11, 25, 384, 256, 5, False, torch.bfloat16, False, True, 0, False, True),    # ragged tiles (125 rows), odd pair count
    (300, 2, 1536, 2560, 1, True, torch.float32, False, False, 0, False, False),  # QKV: fp32 out
    (300, 2, 512, 1536, 1, True, torch.float32, True, False, 0, False, False),    # out-proj: in-place fp32 residual
    (300, 2, 512, 1536, 1, True, torch.float32, True, True, 0, False, False),     # last FFN2: residual + second output
    (300, 2, 256, 3072, 1, True, torch.bfloat16, False, False, 1, False, False),  # FFN1: ReLU, bf16
    (3, 200, 128, 128, 9, False, torch.float32, True, False, 1, True, False),     # activation after the residual
    (1, 130, 128, 384, 1, True, torch.bfloat16, True, False, 2, False, True),     # 2 row tiles, second nearly empty
    (1, 1, 64, 128, 1, True, torch.float32, False, False, 0, False, False),
    (80, 50, 256, 896, 5, False, torch.bfloat16, False, True, 0, False, True),    # column windows, two outputs
    (60, 100, 128, 1152, 1, True, torch.bfloat16, True, False, 2, False, False),  # column windows, residual
]


@pytest.mark.parametrize('rows,K,N,second,act,use_scale', [
    (5000, 64, 768, True, 0, False),      # the stem: only a0 = GELU(BN(conv + b)) is stored
    (4100, 256, 768, False, 1, True),     # one bf16 output, BN affine + bias + ReLU, ragged last row tile
    (6000, 128, 1024, False, 2, False),   # GELU
    (3300, 512, 2560, True, 1, True),
])
def test_gemm_sixteen_warp_epilogue(cuda, rows, K, N, second, act, use_scale):
  """Single-output bf16 launches without residual take the 16-epilogue-warp variant of gemm2_kernel
  (thread = row x column quarter, 16-column TMEM chunks, one slab per quarter) once there are
  enough 256-wide tiles; same reference as the generic epilogue test."""
  g = torch.Generator().manual_seed(rows + K + N)
  A = torch.randn(1, rows, K, generator=g).to(cuda).bfloat16()
  W = (torch.randn(1, N, K, generator=g) * K ** -0.5).to(cuda).bfloat16()
  bias = torch.randn(N, generator=g).to(cuda)
  scale = (1 + 0.2 * torch.randn(N, generator=g)).to(cuda) if use_scale else None
  shift = (0.3 * torch.randn(N, generator=g)).to(cuda) if use_scale else None
  scale2 = (1 + 0.2 * torch.randn(N, generator=g)).to(cuda)
  shift2 = (0.3 * torch.randn(N, generator=g)).to(cuda)
  dst = torch.full((rows + 3, N), 7.0, device=cuda, dtype=torch.bfloat16)     # 3 guard rows
  _lib.selftest_gemm_epilogue(A, W, taps=1, flat=True, bias=bias, scale=scale, shift=shift, act=act,
                              out=None if second else dst, out2=dst if second else None,
                              scale2=scale2 if second else None, shift2=shift2 if second else None,
                              act2=2 if second else 0)
  torch.cuda.synchronize()
  v = A[0].float() @ W[0].float().t()
  if use_scale:
    v = v * scale + shift
  v = _act(v + bias, act)
  if second:
    v = _act(v * scale2 + shift2, 2)
  assert float((dst[:rows].float() - v).abs().max()) < 3e-2 * max(1.0, float(v.abs().max()))
  assert bool((dst[rows:] == 7.0).all()), 'rows past the end were written'


@pytest.mark.parametrize('S,L,K,N,taps', [(7, 100, 256, 256, 5), (40, 50, 128, 896, 5), (3, 25, 64, 128, 5),
                                          (130, 100, 768, 768, 5), (5, 30, 64, 256, 9)])
def test_halo_conv_rows_as_descriptor_offsets(cuda, S, L, K, N, taps):
  """HALO mode: activations laid out as flat rows with taps/2 zero rows either side of every
  sequence; the A tile is fetched once per K block with taps-1 extra rows and each tap is a
  row offset of the UMMA shared-memory descriptor.  Valid rows must equal conv1d with zero
  padding; two outputs (z and GELU(BN(z))) as the Enformer tower uses it."""
  g = torch.Generator().manual_seed(S * 31 + L)
  pad = taps // 2
  Lp = L + 2 * pad
  A = torch.randn(S, L, K, generator=g).bfloat16()
  W = (torch.randn(taps, N, K, generator=g) * (K * taps) ** -0.5).bfloat16()
  bias = torch.randn(N, generator=g)
  scale2 = torch.rand(N, generator=g) + 0.5
  shift2 = torch.randn(N, generator=g) * 0.1
  Ap = torch.zeros(S, Lp, K, dtype=torch.bfloat16)
  Ap[:, pad:pad + L] = A
  out = torch.empty(S * Lp, N, dtype=torch.bfloat16, device=cuda)
  out2 = torch.empty(S * Lp, N, dtype=torch.bfloat16, device=cuda)
  _lib.selftest_gemm_epilogue(Ap.reshape(1, S * Lp, K).to(cuda), W.to(cuda), taps=taps, flat=2, bias=bias.to(cuda),
                              out=out, out2=out2, scale2=scale2.to(cuda), shift2=shift2.to(cuda), act2=2)
  ref = F.conv1d(A.float().transpose(1, 2), W.float().permute(1, 2, 0), bias, padding=pad).transpose(1, 2)
  got = out.float().cpu().reshape(S, Lp, N)[:, pad:pad + L]
  err = float((got - ref).abs().max()) / float(ref.abs().max())
  assert err < 1e-2, err
  v = ref * scale2 + shift2
  ref2 = v * torch.sigmoid(1.702 * v)
  got2 = out2.float().cpu().reshape(S, Lp, N)[:, pad:pad + L]
  err2 = float((got2 - ref2).abs().max()) / float(ref2.abs().max())
  assert err2 < 2e-2, err2


@pytest.mark.parametrize('S,L,K,N,taps,flat,odt,use_res,dual,act,aar,use_scale', EPI_CASES)
def test_gemm_fused_epilogues(cuda, S, L, K, N, taps, flat, odt, use_res, dual, act, aar, use_scale):
  """Every epilogue combination the networks use (bias / BN affine / activation / residual /
  second output, bf16 and fp32, in place) against torch fp32 on the same bf16 operands."""
  g = torch.Generator().manual_seed(S * 17 + L + K)
  A = torch.randn(S, L, K, generator=g).to(cuda).bfloat16()
  W = (torch.randn(taps, N, K, generator=g) * (K * taps) ** -0.5).to(cuda).bfloat16()
  bias = torch.randn(N, generator=g).to(cuda)
  scale = (1 + 0.2 * torch.randn(N, generator=g)).to(cuda) if use_scale else None
  shift = (0.3 * torch.randn(N, generator=g)).to(cuda) if use_scale else None
  scale2 = (1 + 0.2 * torch.randn(N, generator=g)).to(cuda)
  shift2 = (0.3 * torch.randn(N, generator=g)).to(cuda)
  res0 = torch.randn(S * L, N, generator=g).to(cuda).to(odt) if use_res else None
  out = res0.clone() if use_res else torch.full((S * L, N), 7.0, device=cuda, dtype=odt)
  out2 = torch.full((S * L, N), 7.0, device=cuda, dtype=torch.bfloat16 if use_res else odt) if dual else None
  guard = out.clone()
  _lib.selftest_gemm_epilogue(A, W, taps=taps, flat=flat, bias=bias, scale=scale, shift=shift, act=act,
                              act_after_res=aar, res=out if use_res else None, out=out, out2=out2,
                              scale2=scale2 if dual else None, shift2=shift2 if dual else None,
                              act2=2 if dual else 0)
  torch.cuda.synchronize()
  y = F.conv1d(A.float().permute(0, 2, 1), W.float().permute(1, 2, 0), None, padding=taps // 2)
  v = y.permute(0, 2, 1).reshape(S * L, N)
  if use_scale:
    v = v * scale + shift
  v = v + bias
  if not aar:
    v = _act(v, act)
  if use_res:
    v = v + res0.float()
  if aar:
    v = _act(v, act)
  tol = 3e-2 if odt == torch.bfloat16 else 3e-3
  assert float((out.float() - v).abs().max()) < tol * max(1.0, float(v.abs().max()))
  assert not torch.equal(out, guard)
  if dual:
    v2 = _act(v * scale2 + shift2, 2)
    assert float((out2.float() - v2).abs().max()) < 3e-2 * max(1.0, float(v2.abs().max()))
