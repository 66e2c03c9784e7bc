"""Shared test utilities (seeded model builders used by both the golden
generator and the tests, so weights never have to be stored in fixtures)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, GOLDEN):
  if p not in sys.path:
    sys.path.insert(0, p)

from svdd_b200.synthetic import (CONVGRU_HEAD_KW, CONVGRU_ORACLE_KW, CONVGRU_VALUE_KW,  # noqa: E402,F401
                                  ENFORMER_FULL_KW, ENFORMER_SMALL_KW, model_cfg, perturb_,
                                  random_tokens)


from svdd_b200.synthetic import build_convgru_oracle, build_convgru_value, build_denoiser  # noqa: E402,F401
from svdd_b200 import synthetic  # noqa: E402


from svdd_b200.synthetic import build_dit, build_enformer, perturb_dit_, svdd_step_candidates  # noqa: E402,F401


def load_golden(name):
  return np.load(os.path.join(GOLDEN, name))


def state_checksum(sd):
  """Order-independent fingerprint of a state_dict (float64 sums)."""
  tot, tot_abs = 0.0, 0.0
  for k in sorted(sd):
    if sd[k].dtype.is_floating_point:
      v = sd[k].double()
      tot += float(v.sum())
      tot_abs += float(v.abs().sum())
  return np.array([tot, tot_abs])
