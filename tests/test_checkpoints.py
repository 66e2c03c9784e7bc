"""CPU: checkpoint ingestion -- the three artifact layouts the reference's
decode path reads, round-tripped through files:

  * Lightning ``.ckpt`` of the MDLM denoiser   Diffusion.load_from_checkpoint   (Enformer.py:92)
  * value-function ``.pt``                      BaseModel.load_state_dict        (decode.py:101-104,
    written by trainer.py:73-88 as {'model_state_dict': BaseModel.state_dict(), ...})
  * gReLU reward-oracle ``model.ckpt``          load_grelu_reward_model          (Enformer.py:104-131)

No CUDA: loading only moves parameters (the kernels pack them later, on the device)."""
import numpy as np
import pytest
import torch

import helpers
from svdd_b200 import base_model, config, diffusion_gosai, value_nets


def _equal_sd(a, b):
  return sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_lightning_ckpt_round_trip(tmp_path):
  cfg = config.load_config('rna')
  torch.manual_seed(7)
  src = diffusion_gosai.Diffusion(cfg)
  ckpt = {'epoch': 3, 'global_step': 1234, 'pytorch-lightning_version': '2.2.1',
          'state_dict': {k: v.clone() for k, v in src.state_dict().items()},
          'ema': {'decay': 0.9999, 'num_updates': 5, 'shadow_params': [torch.zeros(3)]},
          'optimizer_states': [{}], 'lr_schedulers': [{}], 'hyper_parameters': {'config': 'cfg'}}
  path = tmp_path / 'last.ckpt'
  torch.save(ckpt, path)
  torch.manual_seed(8)                                     # a different init must be overwritten
  dst = diffusion_gosai.Diffusion.load_from_checkpoint(str(path), config=cfg, map_location='cpu')
  assert _equal_sd(src.state_dict(), dst.state_dict())
  with pytest.raises(ValueError):
    diffusion_gosai.Diffusion.load_from_checkpoint(str(path))             # config is required (Enformer.py:92)
  bad = dict(ckpt, state_dict={k: v for k, v in ckpt['state_dict'].items() if 'norms.3.' not in k})
  torch.save(bad, path)
  with pytest.raises(RuntimeError, match='norms.3'):
    diffusion_gosai.Diffusion.load_from_checkpoint(str(path), config=cfg)


def test_lightning_ckpt_written_by_the_reference_class(tmp_path):
  """Build container only: a ``state_dict`` produced by the reference's own Diffusion loads."""
  import ref_import
  if not ref_import.reference_available():
    pytest.skip('reference tree not present')
  ref = ref_import.import_reference()
  torch.manual_seed(11)
  r = ref.diffusion_gosai.Diffusion(ref_import.make_config(length=200))
  path = tmp_path / 'last.ckpt'
  torch.save({'state_dict': r.state_dict(), 'epoch': 0}, path)
  ours = diffusion_gosai.Diffusion.load_from_checkpoint(str(path), config=config.load_config('dna'))
  assert _equal_sd(r.state_dict(), ours.state_dict())


def _reference_layout(model, extra=True):
  """A BaseModel.state_dict() renamed to the reference's layout: the gReLU LightningModel keeps
  its network under ``.model`` (Enformer.py:104-131), Lightning modules carry metric state."""
  sd = {}
  for k, v in model.state_dict().items():
    if k.startswith('reward_model.'):
      k = 'reward_model.model.' + k[len('reward_model.'):]
    sd[k] = v.clone()
  if extra:
    sd['reward_model.val_metrics.mse.sum_squared_error'] = torch.zeros(1)
    sd['reward_model.test_metrics.pearson.n_total'] = torch.zeros(1)
    sd['reward_model.transform.weight'] = torch.zeros(2)
    sd['ref_model.valid_metrics.nll.mean_value'] = torch.zeros(())
  return sd


def _small_dna_reward_model(_task):
  emb = value_nets.EnformerTrunk(n_conv=7, channels=384, n_transformers=1, n_heads=8, key_len=64)
  head = value_nets.ConvHead(n_tasks=3, in_channels=768, act_func=None, pool_func='avg')
  return value_nets.OriBaseModel(emb, head)


@pytest.mark.parametrize('task', ['rna', 'dna'])
def test_value_pt_in_reference_layout_loads_strict(tmp_path, task, monkeypatch):
  """decode.py:101-104: model.load_state_dict(torch.load(path)['model_state_dict'], strict=True)
  with the file trainer.py:73-88 writes.  The reference's keys (reward_model.model.*,
  Lightning metric buffers) are mapped / dropped as documented in BaseModel.map_reference_keys."""
  if task == 'dna':      # full-size nets are 2 GB of fp32: the key logic is size-independent
    monkeypatch.setattr(base_model, 'random_reward_model', _small_dna_reward_model)

  def build(seed):
    torch.manual_seed(seed)
    if task == 'dna':
      emb = value_nets.EnformerTrunk(n_conv=7, channels=384, n_transformers=2, n_heads=8, key_len=64)
      head = value_nets.ConvHead(n_tasks=1, in_channels=768, act_func=None, pool_func='avg')
    else:
      emb = head = None
    return base_model.BaseModel(emb, head, cdq=False, batch_size=4, val_batch_num=1, task=task, random_init=True)

  src, dst = build(1), build(2)
  assert not _equal_sd(src.state_dict(), dst.state_dict())
  path = tmp_path / 'value.pt'
  torch.save({'epoch': 9, 'model_state_dict': _reference_layout(src), 'optimizer_state_dict': {},
              'scaler_state_dict': {}, 'tokens': 0, 'best_loss': 0.5}, path)
  ckpt = torch.load(path, map_location='cpu', weights_only=False)
  dst.load_state_dict(ckpt['model_state_dict'], strict=True)
  assert _equal_sd(src.state_dict(), dst.state_dict())
  assert len(dst.ignored_checkpoint_keys) == 4
  # this package's own layout loads too
  dst2 = build(3)
  dst2.load_state_dict(src.state_dict(), strict=True)
  assert _equal_sd(src.state_dict(), dst2.state_dict())
  # strictness is kept for everything on the path
  sd = _reference_layout(src)
  missing = {k: v for k, v in sd.items() if not k.startswith('head.')}
  with pytest.raises(RuntimeError, match='Missing key'):
    build(4).load_state_dict(missing, strict=True)
  sd['embedding.not_a_parameter'] = torch.zeros(1)
  with pytest.raises(RuntimeError, match='Unexpected key'):
    build(5).load_state_dict(sd, strict=True)


def test_value_pt_from_reference_modules(tmp_path):
  """Build container only: the state_dict pieces come from the reference's own classes
  (EnformerTrunk / ConvHead / Diffusion / OriBaseModel under '.model')."""
  import ref_import
  if not ref_import.reference_available():
    pytest.skip('reference tree not present')
  ref = ref_import.import_reference()
  E = ref.Enformer
  kw = dict(n_conv=7, channels=384, n_transformers=2, n_heads=8, key_len=64)
  torch.manual_seed(21)
  emb, head = E.EnformerTrunk(**kw), E.ConvHead(n_tasks=1, in_channels=768, act_func=None, pool_func='avg')
  den = ref.diffusion_gosai.Diffusion(ref_import.make_config(length=200))
  r_emb = E.EnformerTrunk(n_conv=7, channels=384, n_transformers=1, n_heads=8, key_len=64)
  r_head = E.ConvHead(n_tasks=3, in_channels=768, act_func=None, pool_func='avg')
  sd = {}
  for prefix, mod in (('embedding.', emb), ('head.', head), ('ref_model.', den),
                      ('reward_model.model.embedding.', r_emb), ('reward_model.model.head.', r_head)):
    sd.update({prefix + k: v for k, v in mod.state_dict().items()})
  path = tmp_path / 'value.pt'
  torch.save({'model_state_dict': sd}, path)
  import unittest.mock as mock
  with mock.patch.object(base_model, 'random_reward_model', _small_dna_reward_model):
    ours = base_model.BaseModel(value_nets.EnformerTrunk(**kw),
                                value_nets.ConvHead(n_tasks=1, in_channels=768, act_func=None, pool_func='avg'),
                                cdq=False, batch_size=4, val_batch_num=1, task='dna', random_init=True)
  ours.load_state_dict(torch.load(path, map_location='cpu', weights_only=False)['model_state_dict'], strict=True)
  got = ours.state_dict()
  assert all(torch.equal(got['embedding.' + k], v) for k, v in emb.state_dict().items())
  assert all(torch.equal(got['head.' + k], v) for k, v in head.state_dict().items())
  assert all(torch.equal(got['reward_model.embedding.' + k], v) for k, v in r_emb.state_dict().items())
  assert all(torch.equal(got['reward_model.head.' + k], v) for k, v in r_head.state_dict().items())
  assert all(torch.equal(got['ref_model.' + k], v) for k, v in den.state_dict().items())


def test_grelu_reward_ckpt_round_trip(tmp_path):
  """Enformer.py:104-131 loads the reward oracle with gReLU's LightningModel.load_from_checkpoint:
  a Lightning file whose ``state_dict`` holds ``model.embedding.*`` / ``model.head.*``."""
  torch.manual_seed(31)
  src = base_model.random_reward_model('rna')
  sd = {'model.' + k: v.clone() for k, v in src.state_dict().items()}
  sd['val_metrics.mse.sum_squared_error'] = torch.zeros(1)
  path = tmp_path / 'model.ckpt'
  torch.save({'state_dict': sd, 'hyper_parameters': {'model_params': {}, 'train_params': {}},
              'data_params': {'tasks': {'name': ['MRL']}}}, path)
  torch.manual_seed(32)
  dst = base_model.load_grelu_reward_model(str(path), 'rna')
  assert _equal_sd(src.state_dict(), dst.state_dict())
  del sd['model.head.channel_transform.conv.layer.bias']
  torch.save({'state_dict': sd}, path)
  with pytest.raises(RuntimeError):
    base_model.load_grelu_reward_model(str(path), 'rna')


def test_base_model_reads_artifacts_dir(tmp_path):
  """Without random_init the constructor follows Enformer.py:76-131: hard-coded artifact paths
  (here under artifacts_dir), Lightning denoiser checkpoint + gReLU reward checkpoint."""
  cfg = config.load_config('rna')
  torch.manual_seed(41)
  den = diffusion_gosai.Diffusion(cfg)
  rew = base_model.random_reward_model('rna')
  (tmp_path / 'RNA_Diffusion:v0').mkdir()
  (tmp_path / 'RNA_evaluation:v0').mkdir()
  torch.save({'state_dict': den.state_dict()}, tmp_path / 'RNA_Diffusion:v0' / 'best.ckpt')
  torch.save({'state_dict': {'model.' + k: v for k, v in rew.state_dict().items()}},
             tmp_path / 'RNA_evaluation:v0' / 'model.ckpt')
  torch.manual_seed(42)
  m = base_model.BaseModel(None, None, cdq=False, batch_size=4, val_batch_num=1, task='rna',
                           artifacts_dir=str(tmp_path))
  assert _equal_sd(m.ref_model.state_dict(), den.state_dict())
  assert _equal_sd(m.reward_model.state_dict(), rew.state_dict())
  assert not any(p.requires_grad for p in m.ref_model.parameters())
  with pytest.raises(FileNotFoundError):
    base_model.BaseModel(None, None, cdq=False, batch_size=4, val_batch_num=1, task='dna',
                         artifacts_dir=str(tmp_path))
