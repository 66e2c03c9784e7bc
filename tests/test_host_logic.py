"""CPU: host-side logic -- CLI surface, config, sharding (world_size-2 gloo), npz schema."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers


def test_cli_flags_match_reference_surface():
  import decode
  p = decode.build_parser()
  a = p.parse_args(['--task', 'dna', '--sample_M', '10', '--reward_name', 'HepG2', '--batch_size', '128'])
  assert (a.task, a.sample_M, a.seed, a.val_batch_num, a.model) == ('dna', 10, 44, 1, 'enformer')
  assert p.parse_args([]).task == 'DNA' and p.parse_args([]).sample_M == 5      # decode.py:130,165
  pt = decode.build_parser(tweedie=True)
  with pytest.raises(SystemExit):
    pt.parse_args(['--task', 'rna'])                       # --tweedie is required (decode_tweedie.py:206)
  a = pt.parse_args(['--task', 'rna', '--tweedie', 'True'])
  assert a.tweedie == 'True' and a.sample_M == 20


def test_configs_carry_the_reference_keys():
  from svdd_b200 import config
  for task, L in (('dna', 200), ('DNA', 200), ('rna', 50)):
    c = config.load_config(task)
    assert c.model.length == L and c.model.hidden_dim == 128 and c.model.num_cnn_stacks == 4
    assert c.sampling.predictor == 'ddpm' and c.sampling.steps == 128 and c.sampling.noise_removal
    assert c.parameterization == 'subs' and c.backbone == 'cnn' and c.time_conditioning is False
  if os.path.isdir('/root/reference/configs_gosai'):
    c = config.load_config('dna', '/root/reference/configs_gosai')
    assert c.model.length == 200 and c.sampling.steps == 128


def test_partition_is_a_contiguous_cover():
  from svdd_b200 import sharding
  for B in (0, 1, 7, 128, 4096):
    for ws in (1, 2, 3, 8):
      blocks = [sharding.partition(B, r, ws) for r in range(ws)]
      assert sum(n for _, n in blocks) == B
      pos = 0
      for lo, n in blocks:
        assert lo == pos or n == 0
        pos += n


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from svdd_b200 import sharding
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
B, L = 7, 5
n, off = sharding.local_rows(B)
x = (torch.arange(B * L).reshape(B, L))[off:off + n]        # this rank's rows of a known global tensor
full = sharding.gather_rows(x)
assert full.shape == (B, L) and torch.equal(full, torch.arange(B * L).reshape(B, L)), full
print('rank', dist.get_rank(), 'ok', n, off)
dist.destroy_process_group()
'''


def test_sharded_gather_world_size_2_gloo(tmp_path):
  script = tmp_path / 'worker.py'
  script.write_text(_WORKER)
  procs = []
  for r in range(2):
    env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT='29517')
    procs.append(subprocess.Popen([sys.executable, str(script), helpers.ROOT], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
  outs = [p.communicate(timeout=120)[0].decode() for p in procs]
  assert all(p.returncode == 0 for p in procs), outs
  assert all('ok' in o for o in outs)


def test_npz_schema_of_reference_logs_is_what_we_write(tmp_path):
  """decode.py:117 / decode_tweedie.py:118: keys decoding, baseline; float32 (N,)."""
  ours = np.random.rand(8).astype(np.float32)
  np.savez(tmp_path / 'dna-HepG2', decoding=ours, baseline=ours)
  d = np.load(tmp_path / 'dna-HepG2.npz')
  assert sorted(d.files) == ['baseline', 'decoding'] and d['decoding'].dtype == np.float32
  ref = '/root/reference/log/dna-HepG2.npz'
  if os.path.isfile(ref):
    r = np.load(ref)
    assert sorted(r.files) == sorted(d.files) and r['decoding'].dtype == d['decoding'].dtype
    assert r['decoding'].ndim == d['decoding'].ndim == 1


def test_containers_reproduce_reference_state_dicts():
  """Only where /root/reference exists (build container): same seed -> same tensors."""
  import ref_import
  if not ref_import.reference_available():
    pytest.skip('reference tree not present')
  ref = ref_import.import_reference()
  from svdd_b200 import diffusion_gosai, value_nets, config
  cfg_ref = ref_import.make_config(length=50)
  torch.manual_seed(44)
  a = ref.diffusion_gosai.Diffusion(cfg_ref).state_dict()
  torch.manual_seed(44)
  b = diffusion_gosai.Diffusion(config.load_config('rna')).state_dict()
  assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
  for kw, Ref, Ours in ((helpers.CONVGRU_VALUE_KW, ref.Enformer.ConvGRUTrunk, value_nets.ConvGRUTrunk),
                        (helpers.ENFORMER_SMALL_KW, ref.Enformer.EnformerTrunk, value_nets.EnformerTrunk)):
    torch.manual_seed(7)
    a = Ref(**kw).state_dict()
    torch.manual_seed(7)
    b = Ours(**kw).state_dict()
    assert sorted(a.keys()) == sorted(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
