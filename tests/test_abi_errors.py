"""Error behaviour at the C ABI (include/svdd_b200.h): every entry point returns a negative
status and leaves a message in svdd_last_error() instead of crashing -- the counterpart of the
reference's Python asserts / exceptions on the same misuse (diffusion_gosai.py:1182-1183, 1190;
decode.py:101-104 strict loading).  The argument checks run before any device work, so they are
CPU tests; the ones that need packed weights are marked gpu."""
import ctypes

import pytest
import torch

import helpers
from svdd_b200 import _lib

vp = ctypes.c_void_p
INVALID, CUDA_ERR, UNSUPPORTED, WS_SMALL, MISSING = -1, -2, -3, -4, -5


def _err():
  return _lib.lib().svdd_last_error().decode()


def test_stage2_rejects_bad_shapes_without_touching_a_device():
  L = _lib.lib()
  buf = (ctypes.c_float * 64)()
  tok = (ctypes.c_int64 * 8)()
  p = ctypes.cast(buf, vp)
  t = ctypes.cast(tok, vp)
  # M = 0 candidates
  rc = L.svdd_subs_sample(p, 0, t, _lib.SVDD_TOK_I64, None, 0, None, 0, 0, 0.5, 0.4, t, None, 1, 2, 0, None)
  assert rc == INVALID and 'bad shape' in _err()
  # negative batch
  rc = L.svdd_subs_sample(p, 0, t, _lib.SVDD_TOK_I64, None, 0, None, 0, 0, 0.5, 0.4, t, None, -1, 2, 1, None)
  assert rc == INVALID
  # null logits with a non-empty batch
  rc = L.svdd_subs_sample(None, 0, t, _lib.SVDD_TOK_I64, None, 0, None, 0, 0, 0.5, 0.4, t, None, 1, 2, 1, None)
  assert rc == INVALID and 'null pointer' in _err()
  # unknown token dtype
  rc = L.svdd_subs_sample(p, 0, t, 7, None, 0, None, 0, 0, 0.5, 0.4, t, None, 1, 2, 1, None)
  assert rc == INVALID and 'tok_dtype' in _err()
  # an EMPTY batch is not an error (the reference's loops simply do nothing), even with null pointers
  rc = L.svdd_subs_sample(None, 0, None, _lib.SVDD_TOK_I64, None, 0, None, 0, 0, 0.5, 0.4, None, None, 0, 2, 3, None)
  assert rc == 0


def test_step_rejects_bad_argument_blocks_without_touching_a_device():
  L = _lib.lib()
  assert L.svdd_step(None) == INVALID and 'null argument' in _err()
  a = _lib._StepArgs()
  a.B, a.L, a.M = 2, 50, 3
  assert L.svdd_step(ctypes.byref(a)) == INVALID and 'handles are required' in _err()
  assert L.svdd_step_workspace_bytes(ctypes.byref(a)) == 0
  a.denoiser, a.scorer, a.scorer_kind = 1, 1, 9        # checked before any handle is dereferenced
  assert L.svdd_step(ctypes.byref(a)) == INVALID and 'scorer_kind' in _err()
  a.scorer_kind, a.M = 0, 0
  assert L.svdd_step(ctypes.byref(a)) == INVALID and 'bad shape' in _err()
  a.M, a.tok_dtype = 3, 5
  assert L.svdd_step(ctypes.byref(a)) == INVALID and 'tok_dtype' in _err()
  a.tok_dtype, a.B = _lib.SVDD_TOK_U8, 0               # an empty batch does nothing
  assert L.svdd_step(ctypes.byref(a)) == 0


def test_python_wrappers_raise_with_the_library_message():
  with pytest.raises(_lib.SvddError):           # CPU tensors: no silent fallback
    _lib.subs_sample(torch.zeros(1, 2, 5), torch.zeros(1, 2, dtype=torch.int64), 1, 0.5, 0.4)
  with pytest.raises(_lib.SvddError):
    _lib.select_gather(torch.zeros(1, 2), torch.zeros(2, 1, 3, dtype=torch.int64))


@pytest.mark.gpu
def test_create_names_the_missing_tensor(cuda):
  """decode.py:101-104 loads with strict=True and PyTorch names the missing key; so does svdd_*_create."""
  m = helpers.build_denoiser(44, 50).to(cuda)
  sd = {k: v for k, v in m.state_dict().items() if not k.startswith('time_') and v.dtype.is_floating_point}
  victim = 'convs.3.weight'
  named = [(k, v) for k, v in sd.items() if k != victim]
  arr, keep = _lib._tensor_table(named)
  h = vp()
  rc = _lib.lib().svdd_denoiser_create(arr, len(named), m.num_layers // 5, _lib._stream(), ctypes.byref(h))
  assert rc == MISSING and victim in _err(), _err()
  assert not h.value


@pytest.mark.gpu
def test_forward_checks_workspace_and_arguments(cuda):
  m = helpers.build_denoiser(44, 50).to(cuda)
  den = m.packed()
  x = helpers.random_tokens(4, 50, 5, 0.5).to(cuda)
  tb = den.time_bias(0.0)
  out = torch.empty(4, 50, 5, device=cuda)
  L = _lib.lib()
  need = L.svdd_denoiser_workspace_bytes(den._h, 4, 50)
  ws = torch.empty(max(int(need), 16), dtype=torch.uint8, device=cuda)
  rc = L.svdd_denoiser_forward(den._h, vp(x.data_ptr()), _lib.SVDD_TOK_I64, vp(tb.data_ptr()), vp(out.data_ptr()),
                               4, 50, vp(ws.data_ptr()), 8, _lib._stream())
  assert rc == WS_SMALL and 'workspace too small' in _err(), (rc, _err())
  rc = L.svdd_denoiser_forward(den._h, vp(x.data_ptr()), 9, vp(tb.data_ptr()), vp(out.data_ptr()),
                               4, 50, vp(ws.data_ptr()), ws.numel(), _lib._stream())
  assert rc == INVALID and 'tok_dtype' in _err()
  rc = L.svdd_denoiser_forward(den._h, vp(x.data_ptr()), _lib.SVDD_TOK_I64, vp(tb.data_ptr()), vp(out.data_ptr()),
                               4, 50, vp(ws.data_ptr()), ws.numel(), _lib._stream())
  assert rc == 0, _err()
  torch.cuda.synchronize()
  assert torch.isfinite(out).all()
  # the library is still usable after an error and the handle is intact
  assert torch.equal(den.forward(x, 0.0), den.forward(x, 0.0))
