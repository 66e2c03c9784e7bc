"""Import the UNMODIFIED reference modules from /root/reference on CPU.

Only used by ``tests/golden/make_golden.py`` (fixture generation) and by CPU
tests that skip themselves when /root/reference is absent (it does not exist on
the GPU box).  Third-party packages that the reference imports at module scope
but that are not installed here (lightning, hydra, torchmetrics, timm, grelu,
enformer_pytorch) are replaced by minimal stand-ins in ``sys.modules``;
``enformer_pytorch`` is served by the restatement in ``oracle/enformer_shim.py``.
No reference source is copied into this repository.
"""
import os
import sys
import types

REF_ROOT = os.environ.get('SVDD_REFERENCE_ROOT', '/root/reference')


def reference_available():
  return os.path.isfile(os.path.join(REF_ROOT, 'diffusion_gosai.py'))


def _mod(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  sys.modules[name] = m
  return m


def install_stubs():
  """Idempotently register stand-ins for the absent third-party imports."""
  if 'svdd_ref_stubs_installed' in sys.modules:
    return
  import transformers  # noqa: F401  (must be imported before the stubs)
  import torch
  from torch import nn

  class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
      pass

    @property
    def device(self):
      try:
        return next(self.parameters()).device
      except StopIteration:
        return torch.device('cpu')

    @property
    def dtype(self):
      return torch.float32

    def log(self, *a, **k):
      pass

    def log_dict(self, *a, **k):
      pass

  def rank_zero_only(fn):
    return fn

  L = _mod('lightning', LightningModule=LightningModule)
  Lp = _mod('lightning.pytorch')
  Lu = _mod('lightning.pytorch.utilities', rank_zero_only=rank_zero_only)
  L.pytorch = Lp
  Lp.utilities = Lu

  hydra = _mod('hydra', initialize=lambda *a, **k: None,
               compose=lambda *a, **k: None)
  hu = _mod('hydra.utils', instantiate=lambda *a, **k: None)
  hc = _mod('hydra.core')

  class _GH:
    @staticmethod
    def instance():
      return _GH()

    def clear(self):
      pass

  hg = _mod('hydra.core.global_hydra', GlobalHydra=_GH)
  hydra.utils, hydra.core, hc.global_hydra = hu, hc, hg

  class _Metric(nn.Module):
    def __init__(self, *a, **k):
      super().__init__()

    def update(self, *a, **k):
      pass

    def compute(self):
      return torch.tensor(0.0)

    def reset(self):
      pass

  class MetricCollection(nn.Module):
    def __init__(self, metrics=None, *a, **k):
      super().__init__()

    def set_dtype(self, *a, **k):
      return self

    def clone(self, *a, **k):
      return MetricCollection()

    def update(self, *a, **k):
      pass

    def reset(self):
      pass

  tm = _mod('torchmetrics', MetricCollection=MetricCollection, Metric=_Metric)
  tma = _mod('torchmetrics.aggregation', MeanMetric=_Metric)
  tm.aggregation = tma

  class CosineLRScheduler:
    def __init__(self, *a, **k):
      pass

  timm = _mod('timm')
  ts = _mod('timm.scheduler', CosineLRScheduler=CosineLRScheduler)
  timm.scheduler = ts

  class LightningModel(nn.Module):
    @classmethod
    def load_from_checkpoint(cls, *a, **k):
      raise RuntimeError('grelu checkpoints are unavailable offline')

  grelu = _mod('grelu')
  gl = _mod('grelu.lightning', LightningModel=LightningModel)
  gd = _mod('grelu.data')
  gdp = _mod('grelu.data.preprocess')
  gdd = _mod('grelu.data.dataset')
  grelu.lightning, grelu.data = gl, gd
  gd.preprocess, gd.dataset = gdp, gdd

  # enformer_pytorch <- restated algorithm (oracle/enformer_shim.py)
  repo_root = os.path.dirname(os.path.dirname(os.path.dirname(
      os.path.abspath(__file__))))
  if repo_root not in sys.path:
    sys.path.insert(0, repo_root)
  from oracle import enformer_shim
  ep = _mod('enformer_pytorch')
  epm = _mod('enformer_pytorch.modeling_enformer',
             GELU=enformer_shim.GELU,
             AttentionPool=enformer_shim.AttentionPool,
             relative_shift=enformer_shim.relative_shift,
             Attention=enformer_shim.Attention,
             exponential_linspace_int=enformer_shim.exponential_linspace_int)
  ep.modeling_enformer = epm
  _mod('svdd_ref_stubs_installed')


def import_reference():
  """Returns a namespace with the reference modules
  (diffusion_gosai, dnaconv, noise_schedule, Enformer)."""
  if not reference_available():
    raise RuntimeError(f'reference tree not found at {REF_ROOT}')
  install_stubs()
  import warnings
  # The reference tree is on sys.path only while its modules are imported: it holds top-level
  # names (decode.py, oracle.py ...) that would shadow this repository's own afterwards.
  sys.path.insert(0, REF_ROOT)
  try:
    with warnings.catch_warnings():
      warnings.simplefilter('ignore')
      import noise_schedule
      import models.dnaconv as dnaconv
      import diffusion_gosai
      import Enformer
  finally:
    sys.path.remove(REF_ROOT)
  return types.SimpleNamespace(diffusion_gosai=diffusion_gosai,
                               dnaconv=dnaconv,
                               noise_schedule=noise_schedule,
                               Enformer=Enformer)


def import_reference_dit():
  """models/dit.py of the reference, importable on CPU: omegaconf is stubbed (absent here), and
  the two flash-attn calls (CUDA / Triton only) are replaced by oracle/dit_shim.py's restatements
  inside the imported module.  The reference's ``models/__init__.py`` does not import ``dit``
  (it is commented out), hence the explicit import."""
  import_reference()
  import importlib
  if 'omegaconf' not in sys.modules:
    try:
      import omegaconf  # noqa: F401
    except ImportError:
      _mod('omegaconf', OmegaConf=types.SimpleNamespace(create=lambda d: d))
  from oracle import dit_shim
  sys.path.insert(0, REF_ROOT)
  try:
    dit = importlib.import_module('models.dit')
  finally:
    sys.path.remove(REF_ROOT)
  rot = types.SimpleNamespace(apply_rotary_emb_qkv_=dit_shim.apply_rotary_emb_qkv_)
  dit.flash_attn = types.SimpleNamespace(
      layers=types.SimpleNamespace(rotary=rot),
      flash_attn_interface=types.SimpleNamespace(
          flash_attn_varlen_qkvpacked_func=dit_shim.flash_attn_varlen_qkvpacked_func))
  return dit


def make_config(length=200, hidden_dim=128, num_cnn_stacks=4, steps=128,
                predictor='ddpm', eval_batch_size=8, time_conditioning=False):
  """The subset of the hydra config tree that ``Diffusion`` reads
  (configs_gosai/config_gosai.yaml, configs_gosai/model/dnaconv.yaml)."""
  NS = types.SimpleNamespace
  return NS(
      sampling=NS(predictor=predictor, steps=steps, noise_removal=True),
      eval=NS(gen_ppl_eval_model_name_or_path='gpt2-large'),
      training=NS(antithetic_sampling=True, importance_sampling=False,
                  change_of_variables=False, ema=0, sampling_eps=1e-3),
      parameterization='subs', backbone='cnn',
      model=NS(name='dnaconv', type='cnn', length=length,
               hidden_dim=hidden_dim, num_cnn_stacks=num_cnn_stacks,
               dropout=0.0, clean_data=False, cls_free_guidance=False),
      T=0, subs_masking=False, noise=NS(type='loglinear'),
      optim=NS(lr=3e-4), time_conditioning=time_conditioning,
      loader=NS(eval_batch_size=eval_batch_size))
