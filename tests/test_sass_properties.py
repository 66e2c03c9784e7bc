"""-m "not gpu": static properties of the compiled sm_100a code (cuobjdump -sass of the in-tree
library; no GPU needed).  They pin what round 1 measured to matter:

  * the tensor-core kernels really are tcgen05 + TMA (UTCHMMA / UTMALDG in the SASS);
  * every tcgen05.mma / TMA instruction is issued from an ``elect.sync`` region: inside
    ``if (lane == 0)`` ptxas wraps EACH one in an ELECT / R2UR.BROADCAST loop (c2: 208 -> 218 seq/s);
  * the epilogues of the fused denoiser and of gemm2 reach shared memory with LDS / STS, not
    with generic 16-byte LD.E / ST.E (+ R2UR pairs) (c2: 218.7 -> 221.9 seq/s).
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

import helpers

LIB = os.path.join(helpers.ROOT, 'svdd_b200', 'libsvdd_b200.so')


@pytest.fixture(scope='module')
def sass():
  if shutil.which('cuobjdump') is None or not os.path.exists(LIB):
    pytest.skip('needs cuobjdump and the built library')
  out = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True, timeout=600).stdout
  per = collections.defaultdict(collections.Counter)
  fn = None
  for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      fn = m.group(1)
      continue
    if fn is None:
      continue
    for key, pat in (('umma', 'UTCHMMA'), ('tma_load', 'UTMALDG'), ('bcast', 'R2UR.BROADCAST'),
                     ('generic128', r'\b(LD|ST)\.E\.128 ')):
      if re.search(pat, line):
        per[fn][key] += 1
  return per


def test_tensor_core_kernels_are_tcgen05_and_tma(sass):
  mma = [f for f, c in sass.items() if c['umma']]
  for family in ('den_fused_kernel', 'gemm2_kernel', 'tower_kernel', 'conv_gemm_kernel'):
    ks = [f for f in mma if family in f]
    assert ks, f'no tcgen05.mma in any {family}'
    assert all(sass[f]['tma_load'] > 0 for f in ks), f'{family}: a tensor-core kernel without TMA loads'


def test_async_instructions_are_issued_from_elect_regions(sass):
  # a kernel keeps a few R2UR.BROADCAST in its set-up code; a per-instruction loop would show
  # at least one per UTCHMMA / UTMALDG
  bad = {f: (c['bcast'], c['umma'] + c['tma_load']) for f, c in sass.items()
         if c['umma'] and c['bcast'] >= c['umma'] + c['tma_load']}
  assert not bad, f'ELECT/R2UR.BROADCAST loops around async instructions: {bad}'


def test_epilogues_use_shared_space_accesses(sass):
  # the fused denoiser keeps 16 in its prologue (zeroing of the operand planes, once per launch);
  # before the change it had 240 loads + 64 stores in the per-layer epilogue
  for family, allowed in (('den_fused_kernel', 16), ('gemm2_kernel', 0)):
    bad = {f: c['generic128'] for f, c in sass.items() if family in f and c['generic128'] > allowed}
    assert not bad, f'generic 16-byte LD.E / ST.E in {family}: {bad}'
