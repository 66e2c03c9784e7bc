"""CPU: the C-ABI shared library loads and exports every symbol that
include/svdd_b200.h declares (no compute calls without a GPU), and the product
package never imports the oracle."""
import ctypes
import os
import re

import pytest

import helpers

HEADER = os.path.join(helpers.ROOT, 'include', 'svdd_b200.h')
LIB = os.path.join(helpers.ROOT, 'svdd_b200', 'libsvdd_b200.so')


def _declared_symbols():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(svdd_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_path():
  syms = _declared_symbols()
  for s in ('svdd_subs_sample', 'svdd_select_gather', 'svdd_x0_argmax', 'svdd_denoiser_forward',
            'svdd_convgru_score', 'svdd_enformer_score', 'svdd_last_error', 'svdd_version'):
    assert s in syms


def test_library_exports_every_declared_symbol():
  if not os.path.isfile(LIB):
    import __graft_entry__
    __graft_entry__.build()
  lib = ctypes.CDLL(LIB)
  missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
  assert not missing, missing
  lib.svdd_version.restype = ctypes.c_int
  assert lib.svdd_version() == 100
  lib.svdd_last_error.restype = ctypes.c_char_p
  assert isinstance(lib.svdd_last_error(), bytes)


def test_product_never_imports_the_oracle():
  pkg = os.path.join(helpers.ROOT, 'svdd_b200')
  offenders = []
  for root, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        txt = open(os.path.join(root, f)).read()
        if re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M) or 'oracle/' in txt.replace('oracle/philox.py', ''):
          offenders.append(f)
  for f in ('decode.py', 'decode_tweedie.py'):
    path = os.path.join(helpers.ROOT, f)
    if os.path.isfile(path) and re.search(r'^\s*(from|import)\s+oracle\b', open(path).read(), flags=re.M):
      offenders.append(f)
  assert not offenders, offenders


def test_missing_library_fails_loudly(monkeypatch):
  from svdd_b200 import _lib
  monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libsvdd_b200.so')
  monkeypatch.setattr(_lib, '_lib', None)
  with pytest.raises(_lib.SvddError):
    _lib.lib()
