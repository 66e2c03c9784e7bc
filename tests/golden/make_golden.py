"""Generates the golden fixtures in this directory by running the UNMODIFIED
reference (/root/reference, imported through tests/golden/ref_import.py) on CPU.

    python tests/golden/make_golden.py

The reference tree is only present in the build container, so the outputs are
committed.  Weights are never stored: both the reference modules and the
svdd_b200 containers are built from the same seeds (tests/helpers.py) and a
state_dict checksum in each fixture guards that equivalence.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import helpers  # noqa: E402
import ref_import  # noqa: E402

ref = ref_import.import_reference()
DG = ref.diffusion_gosai
E = ref.Enformer
torch.set_num_threads(8)


def save(name, **arrays):
  path = os.path.join(HERE, name)
  np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                               for k, v in arrays.items()})
  print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


class RandLikeTap:
  """Replaces torch.rand_like inside the reference: injects (or records) the
  uniform tensors of _sample_categorical (diffusion_gosai.py:30-34)."""

  def __init__(self, inject=None):
    self.inject = list(inject) if inject is not None else None
    self.record = []
    self._orig = torch.rand_like

  def __enter__(self):
    def fake(t, *a, **k):
      u = self.inject.pop(0).clone() if self.inject is not None else self._orig(t, *a, **k)
      self.record.append(u.clone())
      return u
    torch.rand_like = fake
    return self

  def __exit__(self, *exc):
    torch.rand_like = self._orig


class FnModule(torch.nn.Module):
  """Lets a plain function stand in for the reference's ``backbone`` module."""

  def __init__(self, fn):
    super().__init__()
    self.fn = fn

  def forward(self, *a, **k):
    return self.fn(*a, **k)


def ref_diffusion(length, seed=44):
  cfg = ref_import.make_config(length=length)
  torch.manual_seed(seed)
  return DG.Diffusion(cfg).eval()


# -- 1. noise schedule ---------------------------------------------------------------
def gen_schedule():
  d = ref_diffusion(50)
  d.backbone = FnModule(lambda x, sigma: torch.zeros(x.shape[0], x.shape[1], 5))
  num_steps, eps = 128, 1e-5
  timesteps = torch.linspace(1, eps, num_steps + 1)
  dt = (1 - eps) / num_steps
  x = torch.full((1, 4), 4, dtype=torch.int64)
  x[0, 1] = 2
  mc_s, diff = [], []
  for i in range(num_steps):
    t = timesteps[i] * torch.ones(1, 1)
    _, _, q, _ = d._ddpm_update_finetune(x, t, dt)
    mc_s.append(q[0, 0, 4].item())
    diff.append(q[0, 1, 2].item())            # exp(0) * (mc_t - mc_s)
  sigma_last = d.noise(timesteps[-1] * torch.ones(1, 1))[0].item()
  save('schedule_128.npz', mc_s=np.float32(mc_s), mc_t_minus_mc_s=np.float32(diff),
       sigma_last=np.float32(sigma_last))


# -- 2. stage-level known answers (injected logits / noise / scores) -------------------
class FakeScorer:
  """Stands in for (embedding, head): records the one-hot candidates it is shown
  and returns injected scores, one column per call."""

  def __init__(self, scores):
    self.scores, self.calls, self.seen = scores, 0, []

  def embedding(self, onehot):
    self.seen.append(onehot.clone())
    return onehot

  def head(self, _):
    s = self.scores[:, self.calls].reshape(-1, 1, 1).clone()
    self.calls += 1
    return s


def onehot_to_tokens(oh):
  tok = oh.argmax(-1)
  return torch.where(oh.sum(-1) == 0, torch.full_like(tok, 4), tok)


def gen_stage_kats():
  d = ref_diffusion(50)
  sched = helpers.load_golden('schedule_128.npz')
  timesteps = torch.linspace(1, 1e-5, 129)
  dt = (1 - 1e-5) / 128
  cases = [('a', 3, 50, 4, 40, 1.0, 0.4), ('b', 2, 200, 10, 5, 1e-6, 0.7),
           ('c', 5, 50, 10, 100, 1.0, 1.0), ('d', 4, 37, 50, 127, 0.05, 0.2)]
  out = {}
  for tag, B, L, M, step, score_scale, p_mask in cases:
    g = torch.Generator().manual_seed(1000 + step)
    logits = torch.randn(B, L, 5, generator=g) * 3
    x = helpers.random_tokens(B, L, 2000 + step, p_mask)
    U = torch.rand(M, B, L, 5, generator=g)
    scores = torch.randn(B, M, generator=g) * score_scale
    if tag == 'c':
      scores[:, 3] = scores[:, 7]              # exact ties -> first index
    d.backbone = FnModule(lambda xx, sigma, lg=logits: lg.clone())
    fake = FakeScorer(scores)
    t = timesteps[step] * torch.ones(B, 1)
    with RandLikeTap(inject=[U[m] for m in range(M)]):
      x_next, _, q, _ = d._ddpm_update_finetune_controlled(
          x, t, dt, fake.embedding, fake.head, repeats=M)
    cand = torch.stack([onehot_to_tokens(o.long()) for o in fake.seen], 0)
    log_p = d.forward(x, torch.zeros(B))
    # noise removal / Tweedie argmax on the same logits
    x0_all = log_p.argmax(dim=2)
    x0_nomask = log_p[:, :, :-1].argmax(dim=-1)
    assert torch.equal(x0_all, x0_nomask)
    for k, v in dict(logits=logits, x=x, U=U, scores=scores, step=step, q=q,
                     cand=cand, x_next=x_next, log_p=log_p, x0=x0_all).items():
      out[f'{tag}_{k}'] = v
    assert abs(q[0, 0, 4].item() - sched['mc_s'][step]) == 0
  save('stage_kats.npz', **out)


# -- 3. denoiser -------------------------------------------------------------------------
def gen_denoiser():
  out = {}
  for L in (50, 200):
    d = ref_diffusion(L)
    x = helpers.random_tokens(3, L, 77 + L, 0.5)
    x[0] = 4                                        # all-mask prior row
    with torch.no_grad():
      logits = d.backbone(x, torch.zeros(3))
      log_p = d.forward(x, torch.zeros(3))
    out[f'L{L}_tokens'], out[f'L{L}_logits'], out[f'L{L}_log_p'] = x, logits, log_p
    out[f'L{L}_checksum'] = helpers.state_checksum(d.backbone.state_dict())
  save('denoiser_seed44.npz', **out)


# -- 4/5. value nets ---------------------------------------------------------------------
def ref_convgru_value():
  torch.manual_seed(3)
  emb = E.ConvGRUTrunk(**helpers.CONVGRU_VALUE_KW)
  head = E.ConvHead(**helpers.CONVGRU_HEAD_KW)
  helpers.perturb_(emb, 7)
  return emb.eval(), head.eval()


def ref_convgru_oracle():
  torch.manual_seed(4)
  emb = E.ConvGRUTrunk(**helpers.CONVGRU_ORACLE_KW)
  head = E.ConvHead(**helpers.CONVGRU_HEAD_KW)
  return E.OriBaseModel(emb, head).eval()


def ref_enformer_small(calibrate=True):
  """Small EnformerTrunk whose BatchNorm running statistics are CALIBRATED on a
  batch (as training would leave them): with the constructor's (0, 1) statistics a
  random-init trunk is so contractive that its output barely depends on the input
  and a parity test could not see semantic errors.  The statistics are committed
  (enformer_small_bn.npz) so the GPU box can rebuild the same net."""
  kw = helpers.ENFORMER_SMALL_KW
  torch.manual_seed(5)
  emb = E.EnformerTrunk(**kw)
  head = E.ConvHead(n_tasks=1, in_channels=2 * kw['channels'], act_func=None, pool_func='avg')
  helpers.perturb_(emb, 9)
  if calibrate:
    bns = [m for m in emb.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    for m in emb.modules():
      if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
    for m in bns:
      m.reset_running_stats()
      m.momentum = None                      # cumulative average over the calibration pass
    emb.train()
    d = ref_diffusion(50)
    cal = helpers.random_tokens(48, 200, 4242, 0.3)
    with torch.no_grad():
      emb(d.transform_samples(cal).float())
    stats = {k: v.clone() for k, v in emb.state_dict().items()
             if k.endswith('running_mean') or k.endswith('running_var')}
    save('enformer_small_bn.npz', **stats)
  return emb.eval(), head.eval()


def gen_value_nets():
  d = ref_diffusion(50)
  out = {}
  with torch.no_grad():
    emb, head = ref_convgru_value()
    tok = helpers.random_tokens(6, 50, 11, 0.3)
    tok[0] = 4
    out['convgru_tokens'] = tok
    out['convgru_values'] = head(emb(d.transform_samples(tok).float())).squeeze()
    out['convgru_checksum'] = helpers.state_checksum(emb.state_dict())
    orc = ref_convgru_oracle()
    out['rnaoracle_values'] = orc(d.transform_samples(tok).float().transpose(1, 2))[:, 0].squeeze()
    out['rnaoracle_checksum'] = helpers.state_checksum(orc.embedding.state_dict())
    emb, head = ref_enformer_small()
    tok = helpers.random_tokens(4, 200, 12, 0.3)
    tok[0] = 4
    out['enformer_tokens'] = tok
    out['enformer_values'] = head(emb(d.transform_samples(tok).float())).squeeze()
    out['enformer_checksum'] = helpers.state_checksum(emb.state_dict())
  save('value_nets.npz', **out)


# -- 6. end-to-end trajectories -----------------------------------------------------------
def gen_trajectories():
  out = {}
  d = ref_diffusion(50)
  emb, head = ref_convgru_value()
  B, M, steps = 4, 3, 12
  torch.manual_seed(123)
  with RandLikeTap() as tap, torch.no_grad():
    x = d.controlled_sample(emb, head, num_steps=steps, eval_sp_size=B, sample_M=M)
  out['mc_tokens'] = x
  out['mc_U'] = torch.stack(tap.record).reshape(steps, M, B, 50, 5)
  orc = ref_convgru_oracle()
  steps = 6
  torch.manual_seed(321)
  with RandLikeTap() as tap, torch.no_grad():
    x = d.controlled_sample_tweedie(orc, num_steps=steps, eval_sp_size=B, sample_M=M,
                                    options='True', task='rna')
  out['pm_tokens'] = x
  out['pm_U'] = torch.stack(tap.record).reshape(steps, M, B, 50, 5)
  torch.manual_seed(55)
  with RandLikeTap() as tap, torch.no_grad():
    x = d.decode_sample(num_steps=16, eval_sp_size=B)
  out['plain_tokens'] = x
  out['plain_U'] = torch.stack(tap.record).reshape(16, 1, B, 50, 5)
  torch.manual_seed(56)
  with torch.no_grad():
    x, mid = d._sample(num_steps=8, eval_sp_size=2)
  out['sample_tokens'] = x
  out['sample_mid'] = torch.stack(mid)
  save('trajectories.npz', **out)


def gen_ddpm_cache():
  """_sample with predictor 'ddpm_cache' (diffusion_gosai.py:755-773, 858-865): many small steps
  so that most steps leave the batch unchanged and reuse the cached p_x0."""
  d = ref_diffusion(50)
  d.sampler = 'ddpm_cache'
  calls = [0]
  orig = d.backbone.forward
  def counted(*a, **k):
    calls[0] += 1
    return orig(*a, **k)
  d.backbone.forward = counted
  B, steps = 2, 160
  torch.manual_seed(77)
  with RandLikeTap() as tap, torch.no_grad():
    x, mid = d._sample(num_steps=steps, eval_sp_size=B)
  save('ddpm_cache.npz', tokens=x, mid=torch.stack(mid), U=torch.stack(tap.record).reshape(steps, 1, B, 50, 5),
       n_forward=np.int64(calls[0]))


if __name__ == '__main__':
  gen_schedule()
  gen_stage_kats()
  gen_denoiser()
  gen_value_nets()
  gen_trajectories()
  gen_ddpm_cache()
