"""ctypes binding of libsvdd_b200.so (the C ABI in include/svdd_b200.h).

PyTorch is used here only as plumbing: device allocations, the current CUDA
stream and data pointers.  Every compute call is a hand-written sm_100a kernel
behind the C ABI.  There is NO fallback: if the shared library has not been
built (``make`` / ``__graft_entry__.build()``) or a tensor is not on a CUDA
device, the call raises.
"""
import ctypes
import itertools
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVDD_LIB_PATH: an alternative build of the same library (A/B experiments)
LIB_PATH = os.environ.get('SVDD_LIB_PATH') or os.path.join(_HERE, 'libsvdd_b200.so')

SVDD_TOK_I64, SVDD_TOK_U8 = 0, 1

_lib = None
_lock = threading.Lock()


class SvddError(RuntimeError):
  pass


class _Tensor(ctypes.Structure):
  _fields_ = [('name', ctypes.c_char_p), ('data', ctypes.c_void_p),
              ('ndim', ctypes.c_int32), ('shape', ctypes.c_int64 * 4)]


class _StepArgs(ctypes.Structure):
  """svdd_step_args (include/svdd_b200.h), field for field."""
  _fields_ = [('denoiser', ctypes.c_void_p), ('time_bias_t', ctypes.c_void_p),
              ('time_bias_s', ctypes.c_void_p), ('scorer_kind', ctypes.c_int),
              ('scorer', ctypes.c_void_p), ('tweedie', ctypes.c_int), ('x', ctypes.c_void_p),
              ('x_out', ctypes.c_void_p), ('tok_dtype', ctypes.c_int), ('idx_out', ctypes.c_void_p),
              ('B', ctypes.c_int), ('L', ctypes.c_int), ('M', ctypes.c_int),
              ('mc_t', ctypes.c_float), ('mc_s', ctypes.c_float), ('alpha', ctypes.c_float),
              ('U', ctypes.c_void_p), ('U_sel', ctypes.c_void_p), ('seed', ctypes.c_uint64),
              ('seed_dev', ctypes.c_void_p), ('step', ctypes.c_int), ('row_offset', ctypes.c_int64),
              ('cand_out', ctypes.c_void_p), ('scores_out', ctypes.c_void_p),
              ('q_out', ctypes.c_void_p), ('ws', ctypes.c_void_p), ('ws_bytes', ctypes.c_size_t),
              ('stream', ctypes.c_void_p)]


def _declare(lib):
  c = ctypes
  vp, i32, i64, u64, f32 = c.c_void_p, c.c_int, c.c_int64, c.c_uint64, c.c_float
  lib.svdd_version.restype = i32
  lib.svdd_last_error.restype = c.c_char_p
  lib.svdd_device_check.argtypes = [i32]
  lib.svdd_launch_count.restype = i64
  lib.svdd_subs_sample.argtypes = [vp, i32, vp, i32, vp, u64, vp, i32, i64, f32, f32,
                                   vp, vp, i32, i32, i32, vp]
  lib.svdd_selftest_subs_sample_exact.argtypes = lib.svdd_subs_sample.argtypes
  lib.svdd_select_gather.argtypes = [vp, vp, i32, f32, vp, u64, vp, i32, i64, vp, vp,
                                     i32, i32, i32, vp]
  lib.svdd_x0_argmax.argtypes = [vp, vp, i32, vp, i64, i32, vp]
  lib.svdd_subs_log_p.argtypes = [vp, vp, i32, vp, i64, i32, vp]
  lib.svdd_selftest_conv_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                          i32, i32, vp]
  lib.svdd_selftest_gemm_epilogue.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, i32, vp, i32, vp, i32,
                                              vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
  lib.svdd_selftest_pool.argtypes = [vp, vp, vp, i32, i32, i32, vp]
  lib.svdd_selftest_pair_pool.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
  lib.svdd_selftest_rel_positions.argtypes = [i32, i32, vp]
  lib.svdd_selftest_attention.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
  tp = c.POINTER(_Tensor)
  for net, extra in (('denoiser', [i32]), ('convgru', []), ('enformer', [i32]), ('dit', [i32])):
    create = getattr(lib, f'svdd_{net}_create', None)
    if create is None:
      continue
    create.argtypes = [tp, i32] + extra + [vp, c.POINTER(vp)]
    getattr(lib, f'svdd_{net}_destroy').argtypes = [vp]
    getattr(lib, f'svdd_{net}_destroy').restype = None
    ws = getattr(lib, f'svdd_{net}_workspace_bytes')
    ws.argtypes = [vp, i64, i32]
    ws.restype = c.c_size_t
  if hasattr(lib, 'svdd_denoiser_forward'):
    lib.svdd_denoiser_forward.argtypes = [vp, vp, i32, vp, vp, i64, i32, vp,
                                          c.c_size_t, vp]
  if hasattr(lib, 'svdd_dit_forward'):
    lib.svdd_dit_forward.argtypes = [vp, vp, i32, vp, vp, i64, i32, vp, c.c_size_t, vp]
    lib.svdd_dit_mod_floats.argtypes = [vp]
    lib.svdd_dit_mod_floats.restype = i64
  for net in ('convgru', 'enformer'):
    fn = getattr(lib, f'svdd_{net}_score', None)
    if fn is not None:
      fn.argtypes = [vp, vp, i32, vp, i64, i32, vp, c.c_size_t, vp]
  lib.svdd_step.argtypes = [c.POINTER(_StepArgs)]
  lib.svdd_step_workspace_bytes.argtypes = [c.POINTER(_StepArgs)]
  lib.svdd_step_workspace_bytes.restype = c.c_size_t


def lib():
  """Loads the shared library or raises (no CPU fallback)."""
  global _lib
  if _lib is None:
    with _lock:
      if _lib is None:
        if not os.path.isfile(LIB_PATH):
          raise SvddError(
              f'{LIB_PATH} is missing: build it with `make` (or '
              '`python -c "import __graft_entry__ as g; g.build()"`). '
              'svdd_b200 has no CPU / PyTorch fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
  return _lib


def check(rc):
  if rc != 0:
    msg = lib().svdd_last_error()
    raise SvddError(f'libsvdd_b200 error {rc}: {msg.decode() if msg else ""}')


def launch_count():
  return int(lib().svdd_launch_count())


def profile_begin():
  check(lib().svdd_profile_begin())


def profile_top():
  """Longest GEMM launch of the last profile window: dict(ms, flops, rows, K, N, taps, mode)."""
  ms, fl = ctypes.c_double(), ctypes.c_double()
  rows = ctypes.c_int64()
  K, N, taps, mode = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
  check(lib().svdd_profile_top(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(rows), ctypes.byref(K),
                               ctypes.byref(N), ctypes.byref(taps), ctypes.byref(mode)))
  return dict(ms=ms.value, flops=fl.value, rows=rows.value, K=K.value, N=N.value, taps=taps.value, mode=mode.value)


def profile_end():
  """-> (ms summed over conv_gemm launches, number of launches, nominal dense FLOPs)."""
  ms, n, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
  check(lib().svdd_profile_end(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)))
  return ms.value, n.value, fl.value


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise SvddError('svdd_b200 kernels need CUDA tensors (there is no CPU path)')


def _ptr(t):
  return ctypes.c_void_p(0 if t is None else t.data_ptr())


def tok_dtype(t):
  if t.dtype == torch.int64:
    return SVDD_TOK_I64
  if t.dtype == torch.uint8:
    return SVDD_TOK_U8
  raise SvddError(f'token tensors must be int64 or uint8, got {t.dtype}')


# -- stage 2 ---------------------------------------------------------------------
def subs_sample(logits, x, M, mc_t, mc_s, U=None, seed=0, step=0, row_offset=0,
                is_log_p=False, want_q=False, out=None, seed_dev=None, exact=False):
  """svdd_subs_sample: logits fp32[B,L,5], x [B,L] -> candidates [M,B,L]
  (+ q_xs fp32[B,L,5] when want_q).  exact=True is the test hook that forces every draw
  onto the exact-arithmetic path (svdd_selftest_subs_sample_exact)."""
  _require_cuda(logits, x, U)
  B, L = x.shape
  assert logits.shape == (B, L, 5) and logits.dtype == torch.float32
  logits, x = logits.contiguous(), x.contiguous()
  if U is not None:
    assert U.shape == (M, B, L, 5) and U.dtype == torch.float32
    U = U.contiguous()
  cand = out if out is not None else torch.empty((M, B, L), dtype=x.dtype, device=x.device)
  q = torch.empty((B, L, 5), dtype=torch.float32, device=x.device) if want_q else None
  fn = lib().svdd_selftest_subs_sample_exact if exact else lib().svdd_subs_sample
  check(fn(_ptr(logits), int(is_log_p), _ptr(x), tok_dtype(x), _ptr(U), int(seed),
           _ptr(seed_dev), int(step), int(row_offset), float(mc_t), float(mc_s), _ptr(cand),
           _ptr(q), B, L, M, _stream()))
  return (cand, q) if want_q else cand


# -- stage 4 ---------------------------------------------------------------------
def select_gather(scores, cand, alpha=0.0, U_sel=None, seed=0, step=0,
                  row_offset=0, want_idx=False, out=None, seed_dev=None):
  """svdd_select_gather: scores fp32[M,B], cand [M,B,L] -> x_next [B,L]."""
  _require_cuda(scores, cand, U_sel)
  M, B, L = cand.shape
  assert scores.shape == (M, B) and scores.dtype == torch.float32
  scores, cand = scores.contiguous(), cand.contiguous()
  if U_sel is not None:
    assert U_sel.shape == (B, M) and U_sel.dtype == torch.float32
    U_sel = U_sel.contiguous()
  x_out = out if out is not None else torch.empty((B, L), dtype=cand.dtype, device=cand.device)
  idx = torch.empty((B,), dtype=torch.int32, device=cand.device) if want_idx else None
  check(lib().svdd_select_gather(_ptr(scores), _ptr(cand), tok_dtype(cand), float(alpha),
                                 _ptr(U_sel), int(seed), _ptr(seed_dev), int(step),
                                 int(row_offset), _ptr(x_out), _ptr(idx), B, L, M, _stream()))
  return (x_out, idx) if want_idx else x_out


# -- one whole reverse step ----------------------------------------------------------
def step(denoiser, scorer, x, M, mc_t, mc_s, sigma_t=0.0, sigma_s=None, tweedie=False, alpha=0.0,
         U=None, U_sel=None, seed=0, step=0, row_offset=0, seed_dev=None, want_q=False, out=None, ws=None):
  """svdd_step: stages 1-4 of one controlled reverse step behind ONE C call
  (Diffusion._ddpm_update_finetune_controlled / _controlled_twedie, diffusion_gosai.py:1175-1228 /
  1374-1460).  denoiser: DenoiserHandle; scorer: ConvGRUHandle / EnformerHandle (the value net for
  MC, the reward oracle for PM).  -> dict(x, idx, cand, scores, q)."""
  _require_cuda(x, U, U_sel)
  x = x.contiguous()
  B, L = x.shape
  dev = x.device
  a = _StepArgs()
  tb_t = denoiser.time_bias(sigma_t)
  tb_s = denoiser.time_bias(sigma_t if sigma_s is None else sigma_s)
  a.denoiser, a.time_bias_t, a.time_bias_s = denoiser._h, tb_t.data_ptr(), tb_s.data_ptr()
  a.scorer_kind = {'convgru': 0, 'enformer': 1}[scorer._net]
  a.scorer, a.tweedie = scorer._h, int(bool(tweedie))
  x_out = out if out is not None else torch.empty_like(x)
  idx = torch.empty((B,), dtype=torch.int32, device=dev)
  cand = torch.empty((M, B, L), dtype=x.dtype, device=dev)
  scores = torch.empty((M, B), dtype=torch.float32, device=dev)
  q = torch.empty((B, L, 5), dtype=torch.float32, device=dev) if want_q else None
  a.x, a.x_out, a.tok_dtype, a.idx_out = x.data_ptr(), x_out.data_ptr(), tok_dtype(x), idx.data_ptr()
  a.B, a.L, a.M = B, L, M
  a.mc_t, a.mc_s, a.alpha = float(mc_t), float(mc_s), float(alpha)
  if U is not None:
    assert U.shape == (M, B, L, 5) and U.dtype == torch.float32
    U = U.contiguous()
    a.U = U.data_ptr()
  if U_sel is not None:
    assert U_sel.shape == (B, M) and U_sel.dtype == torch.float32
    U_sel = U_sel.contiguous()
    a.U_sel = U_sel.data_ptr()
  a.seed, a.step, a.row_offset = int(seed), int(step), int(row_offset)
  a.seed_dev = None if seed_dev is None else seed_dev.data_ptr()
  a.cand_out, a.scores_out = cand.data_ptr(), scores.data_ptr()
  a.q_out = None if q is None else q.data_ptr()
  need = int(lib().svdd_step_workspace_bytes(ctypes.byref(a)))
  if need == 0 and B > 0:
    check(-1)
  if ws is None or ws.numel() < need:
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
  a.ws, a.ws_bytes, a.stream = ws.data_ptr(), need, _stream().value
  check(lib().svdd_step(ctypes.byref(a)))
  return dict(x=x_out, idx=idx, cand=cand, scores=scores, q=q, ws=ws)


def x0_argmax(logits, x, out=None):
  """svdd_x0_argmax: argmax over the 4 real tokens of SUBS(logits, x)."""
  _require_cuda(logits, x)
  L = x.shape[-1]
  n = x.numel() // L
  logits, x = logits.contiguous(), x.contiguous()
  out = out if out is not None else torch.empty_like(x)
  check(lib().svdd_x0_argmax(_ptr(logits), _ptr(x), tok_dtype(x), _ptr(out), n, L, _stream()))
  return out


def subs_log_p(logits, x):
  """svdd_subs_log_p: raw logits + tokens -> post-SUBS log-probs (Diffusion.forward)."""
  _require_cuda(logits, x)
  L = x.shape[-1]
  n = x.numel() // L
  logits, x = logits.contiguous(), x.contiguous()
  out = torch.empty_like(logits)
  check(lib().svdd_subs_log_p(_ptr(logits), _ptr(x), tok_dtype(x), _ptr(out), n, L, _stream()))
  return out


def selftest_conv_gemm(A, W, bias, taps, dil, tensor_cores=True):
  """A bf16[S,L,K], W bf16[taps,N,K], bias fp32[N] -> fp32[S*L,N]."""
  _require_cuda(A, W, bias)
  S, L, K = A.shape
  N = W.shape[1]
  assert W.shape == (taps, N, K) and A.dtype == W.dtype == torch.bfloat16
  C = torch.empty((S * L, N), dtype=torch.float32, device=A.device)
  check(lib().svdd_selftest_conv_gemm(_ptr(A.contiguous()), _ptr(W.contiguous()), _ptr(bias),
                                      _ptr(C), S, L, K, N, taps, dil, int(tensor_cores),
                                      _stream()))
  return C


_DT = {torch.bfloat16: 1, torch.float32: 2}


def selftest_gemm_epilogue(A, W, taps=1, dil=1, flat=False, bias=None, scale=None, shift=None, act=0,
                           act_after_res=False, res=None, out=None, out2=None, scale2=None,
                           shift2=None, act2=0):
  """EPI_GENERIC chain on A bf16[S,L,K], W bf16[taps,N,K]; writes `out` / `out2` ([S*L,N]
  bf16 or fp32 tensors supplied by the caller; `res` may alias `out`)."""
  _require_cuda(A, W, bias, scale, shift, res, out, out2, scale2, shift2)
  S, L, K = A.shape
  N = W.shape[1]
  dt = lambda t: 0 if t is None else _DT[t.dtype]
  check(lib().svdd_selftest_gemm_epilogue(
      _ptr(A.contiguous()), _ptr(W.contiguous()), _ptr(bias), _ptr(scale), _ptr(shift), int(act),
      int(act_after_res), _ptr(res), dt(res), _ptr(out), dt(out), _ptr(out2), dt(out2), _ptr(scale2),
      _ptr(shift2), int(act2), S, L, K, N, taps, dil, int(flat), _stream()))


def selftest_pool(y, Wp):
  """y bf16[S,L,C], Wp bf16[C,C] -> attention-pooled fp32 [S*ceil(L/2), C]."""
  _require_cuda(y, Wp)
  S, L, C = y.shape
  out = torch.empty((S * ((L + 1) // 2), C), dtype=torch.float32, device=y.device)
  pad = torch.zeros((S * L + 8, C), dtype=torch.bfloat16, device=y.device)   # 1 row of slack
  pad[:S * L] = y.reshape(S * L, C)
  check(lib().svdd_selftest_pool(_ptr(pad), _ptr(Wp.contiguous()), _ptr(out), S, L, C, _stream()))
  return out


def selftest_pair_pool(A, W1, bias, res, Wp, scale2=None, shift2=None, want_act=False):
  """Pair-split 1x1 conv + difference pooling (EPI_PAIR -> EPI_POOL2).  A, res bf16 [S,L,C];
  W1, Wp bf16 [C,C] -> (y0, yd bf16 [S,Lo,C], pooled fp32 [S*Lo,C], pooled_act bf16 or None)."""
  _require_cuda(A, W1, res, Wp)
  S, L, C = A.shape
  Lo = (L + 1) // 2
  dev = A.device
  y0 = torch.empty((S, Lo, C), dtype=torch.bfloat16, device=dev)
  yd = torch.empty((S, Lo, C), dtype=torch.bfloat16, device=dev)
  pooled = torch.empty((S * Lo, C), dtype=torch.float32, device=dev)
  pact = torch.empty((S * Lo, C), dtype=torch.bfloat16, device=dev) if want_act else None
  check(lib().svdd_selftest_pair_pool(_ptr(A.contiguous()), _ptr(W1.contiguous()), _ptr(bias),
                                      _ptr(res.contiguous()), _ptr(Wp.contiguous()), _ptr(y0), _ptr(yd),
                                      _ptr(pooled), _ptr(pact), _ptr(scale2), _ptr(shift2), S, L, C,
                                      _stream()))
  return y0, yd, pooled, pact


def selftest_rel_positions(n, F):
  out = torch.empty((2 * n - 1, F), dtype=torch.float32)
  check(lib().svdd_selftest_rel_positions(n, F, ctypes.c_void_p(out.data_ptr())))
  return out


def selftest_attention(qkv, rcb, rpb, relk, n, H, dk, dv):
  _require_cuda(qkv, rcb, rpb, relk)
  rows = qkv.shape[0] // n
  out = torch.empty((rows * n, H * dv), dtype=torch.bfloat16, device=qkv.device)
  check(lib().svdd_selftest_attention(_ptr(qkv.contiguous()), _ptr(rcb.contiguous()),
                                      _ptr(rpb.contiguous()), _ptr(relk.contiguous()), _ptr(out),
                                      rows, n, H, dk, dv, _stream()))
  return out


# -- network handles ---------------------------------------------------------------
def _tensor_table(named):
  """[(name, fp32 CUDA tensor)] -> (ctypes array, keepalive list)."""
  arr = (_Tensor * len(named))()
  keep = []
  for i, (name, t) in enumerate(named):
    t = t.detach().to(torch.float32).contiguous()
    _require_cuda(t)
    keep.append(t)
    arr[i].name = name.encode()
    arr[i].data = t.data_ptr()
    arr[i].ndim = t.dim()
    for d in range(4):
      arr[i].shape[d] = t.shape[d] if d < t.dim() else 1
  return arr, keep


class _Handle:
  """Owns an opaque svdd_* handle plus a grow-only workspace.

  ``uid`` is unique per handle for the life of the process (cache keys must not use ``id()``,
  which is recycled).  A captured CUDA graph keeps raw pointers into the workspace and the
  packed weights: whoever captures must hold ``keepalive()`` for as long as the graph lives
  (Diffusion._graph_trajectory does)."""
  _net = None
  _uids = itertools.count(1)

  def __init__(self):
    self._h = ctypes.c_void_p()
    self._ws = None
    self.uid = next(_Handle._uids)

  def keepalive(self):
    """Objects a captured graph must pin: the handle itself (its packed device weights are freed
    by __del__) and the CURRENT workspace tensor (replaced, never resized in place, when a later
    call needs more rows)."""
    return [self, self._ws]

  def _create(self, named, *extra):
    arr, keep = _tensor_table(named)
    fn = getattr(lib(), f'svdd_{self._net}_create')
    check(fn(arr, len(named), *extra, _stream(), ctypes.byref(self._h)))
    torch.cuda.current_stream().synchronize()   # packing reads `keep`
    del keep

  def _workspace(self, n_rows, L, device):
    need = int(getattr(lib(), f'svdd_{self._net}_workspace_bytes')(self._h, n_rows, L))
    if self._ws is None or self._ws.numel() < need or self._ws.device != device:
      self._ws = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
    return self._ws, need

  def __del__(self):
    try:
      if self._h:
        getattr(lib(), f'svdd_{self._net}_destroy')(self._h)
        self._h = ctypes.c_void_p()
    except Exception:
      pass


class DenoiserHandle(_Handle):
  """Packed CNN denoiser (svdd_denoiser_*).  Built from a
  ``svdd_b200.denoiser.CNNModel`` (or any module with the reference's
  ``CNNModel`` state_dict layout plus ``time_bias``)."""
  _net = 'denoiser'

  def __init__(self, module):
    super().__init__()
    sd = module.state_dict()
    named = [(k, v) for k, v in sd.items()
             if not k.startswith('time_') and v.dtype.is_floating_point]
    dev = next(iter(sd.values())).device
    if dev.type != 'cuda':
      raise SvddError('move the model to a CUDA device before running it')
    stacks = module.num_layers // 5
    self._create(named, stacks)
    self._module = module
    self._tbias_cache = {}

  def time_bias(self, sigma):
    """Per-layer time-conditioning rows for one sigma (host-side nn.Linear's at pack time, cached).
    Must not miss inside a CUDA-graph capture (it would run a pageable H2D copy and cuBLAS):
    Diffusion._graph_trajectory pre-computes every sigma of the schedule first and pins the
    tensors with the graph."""
    key = float(sigma)
    tb = self._tbias_cache.get(key)
    if tb is None:
      if torch.cuda.is_current_stream_capturing():
        raise SvddError('time_bias cache miss during CUDA-graph capture (sigma=%r)' % key)
      tb = self._module.time_bias(key).float().contiguous()
      if len(self._tbias_cache) > 4096:
        self._tbias_cache.clear()       # tensors pinned by captured graphs stay alive through them
      self._tbias_cache[key] = tb
    return tb

  def forward(self, tokens, sigma=0.0, out=None):
    """tokens int64/uint8 [N,L] (CUDA) -> raw logits fp32 [N,L,5]."""
    _require_cuda(tokens)
    tokens = tokens.contiguous()
    N, L = tokens.shape
    tb = self.time_bias(sigma)
    logits = out if out is not None else torch.empty((N, L, 5), dtype=torch.float32,
                                                     device=tokens.device)
    ws, need = self._workspace(N, L, tokens.device)
    check(lib().svdd_denoiser_forward(self._h, _ptr(tokens), tok_dtype(tokens), _ptr(tb),
                                      _ptr(logits), N, L, _ptr(ws), need, _stream()))
    return logits


class DiTHandle(_Handle):
  """Packed DiT denoiser (svdd_dit_*), built from a ``svdd_b200.dit.DIT``.  Same call shape as
  DenoiserHandle: ``forward(tokens, sigma, out)`` -> raw logits fp32 [N,L,5]; the adaLN
  modulation of ``sigma`` is evaluated by the module on the host (cached per sigma) and handed
  to the kernels as folded per-channel vectors."""
  _net = 'dit'

  def __init__(self, module):
    super().__init__()
    sd = module.state_dict()
    skip = ('sigma_map.', 'adaLN_modulation', 'norm1.', 'norm2.', 'norm_final.', 'mlp.2.bias')
    named = [(k, v) for k, v in sd.items()
             if v.dtype.is_floating_point and not any(s in k for s in skip)]
    dev = next(iter(sd.values())).device
    if dev.type != 'cuda':
      raise SvddError('move the model to a CUDA device before running it')
    self._create(named, int(module.n_heads))
    self._module = module
    self._tbias_cache = {}
    self.mod_floats = int(lib().svdd_dit_mod_floats(self._h))

  def time_bias(self, sigma):
    """The folded modulation vectors for one sigma (cached; must not miss inside a graph capture)."""
    key = float(sigma)
    tb = self._tbias_cache.get(key)
    if tb is None:
      if torch.cuda.is_current_stream_capturing():
        raise SvddError('DiT modulation cache miss during CUDA-graph capture (sigma=%r)' % key)
      tb = self._module.modulation(key).float().contiguous()
      assert tb.numel() == self.mod_floats, (tb.numel(), self.mod_floats)
      if len(self._tbias_cache) > 4096:
        self._tbias_cache.clear()
      self._tbias_cache[key] = tb
    return tb

  def forward(self, tokens, sigma=0.0, out=None):
    """tokens int64/uint8 [N,L] (CUDA) -> raw logits fp32 [N,L,5]."""
    _require_cuda(tokens)
    tokens = tokens.contiguous()
    N, L = tokens.shape
    mod = self.time_bias(sigma)
    logits = out if out is not None else torch.empty((N, L, 5), dtype=torch.float32, device=tokens.device)
    ws, need = self._workspace(N, L, tokens.device)
    check(lib().svdd_dit_forward(self._h, _ptr(tokens), tok_dtype(tokens), _ptr(mod), _ptr(logits), N, L,
                                 _ptr(ws), need, _stream()))
    return logits


class _ScoreHandle(_Handle):
  def score(self, tokens, out=None):
    """tokens int64/uint8 [N,L] (CUDA; 4 = mask) -> fp32 [N]."""
    _require_cuda(tokens)
    tokens = tokens.contiguous()
    L = tokens.shape[-1]
    N = tokens.numel() // L
    scores = out if out is not None else torch.empty((N,), dtype=torch.float32,
                                                     device=tokens.device)
    ws, need = self._workspace(N, L, tokens.device)
    fn = getattr(lib(), f'svdd_{self._net}_score')
    check(fn(self._h, _ptr(tokens), tok_dtype(tokens), _ptr(scores), N, L, _ptr(ws), need,
             _stream()))
    return scores


def _named_with_head(sd_embedding, sd_head, task=0):
  """Trunk tensors under their own names + the head's under 'head.'.  A multi-task head
  (the DNA oracle has 3) contributes ONE task, by default task 0: the path reads
  ``reward_model(x)[:, 0]`` (diffusion_gosai.py:1430, Enformer.py:447)."""
  named = [(k, v) for k, v in sd_embedding.items() if v.dtype.is_floating_point]
  named += [('head.' + k, v[task:task + 1]) for k, v in sd_head.items() if v.dtype.is_floating_point]
  if not named or named[0][1].device.type != 'cuda':
    raise SvddError('move the value network to a CUDA device before scoring')
  return named


class ConvGRUHandle(_ScoreHandle):
  _net = 'convgru'

  def __init__(self, sd_embedding, sd_head, task=0):
    super().__init__()
    self._create(_named_with_head(sd_embedding, sd_head, task))


class EnformerHandle(_ScoreHandle):
  _net = 'enformer'

  def __init__(self, sd_embedding, sd_head, n_heads=8, task=0):
    super().__init__()
    self._create(_named_with_head(sd_embedding, sd_head, task), int(n_heads))
