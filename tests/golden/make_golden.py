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


DATA = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'svdd_b200', 'data')


def ref_enformer(full=False, calibrate=True, n_tasks=1):
  """Reference EnformerTrunk (small: 384 ch / 2 blocks; full: decode.py:78-80's 1536 ch / 11
  blocks) whose BatchNorm running statistics are CALIBRATED on a batch (as training would leave
  them): with the constructor's (0, 1) statistics a random-init trunk is so contractive that
  its output barely depends on the input and a parity test could not see semantic errors.  The
  statistics are committed (svdd_b200/data/enformer_{small,full}_bn.npz) so that the GPU box,
  bench.py and the tests rebuild the same net.  calibrate=False reloads the committed ones."""
  kw = helpers.ENFORMER_FULL_KW if full else helpers.ENFORMER_SMALL_KW
  name = 'enformer_full_bn.npz' if full else 'enformer_small_bn.npz'
  torch.manual_seed(5)
  emb = E.EnformerTrunk(**kw)
  head = E.ConvHead(n_tasks=n_tasks, in_channels=2 * kw['channels'], act_func=None, pool_func='avg')
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
    path = os.path.join(DATA, name)
    np.savez_compressed(path, **{k: v.numpy() for k, v in stats.items()})
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')
  else:
    stats = np.load(os.path.join(DATA, name))
    sd = emb.state_dict()
    with torch.no_grad():
      for k in stats.files:
        sd[k].copy_(torch.from_numpy(stats[k]))
  return emb.eval(), head.eval()


def ref_enformer_small(calibrate=True):
  return ref_enformer(False, calibrate)


class ChannelsFirstTrunk(torch.nn.Module):
  """The DNA reward oracle is gReLU's own EnformerModel (Enformer.py:104-131, oracle.py:72), whose
  trunk consumes [N, 4, L] -- which is why the path calls reward_model(onehot.transpose(1, 2))
  (diffusion_gosai.py:1430).  The reference's in-tree EnformerTrunk is the same network with a
  transpose prepended (Enformer.py:1328), so the stand-in un-transposes in front of it."""

  def __init__(self, trunk):
    super().__init__()
    self.trunk = trunk

  def forward(self, x):
    return self.trunk(x.transpose(1, 2))


def ref_dna_reward_model(emb, seed=6):
  """OriBaseModel(trunk [N,4,L] -> [N,3072,2], ConvHead(n_tasks=3)): output [N, 3, 1]."""
  torch.manual_seed(seed)
  head3 = E.ConvHead(n_tasks=3, in_channels=2 * 1536, act_func=None, pool_func='avg').eval()
  return E.OriBaseModel(ChannelsFirstTrunk(emb), head3).eval()


def gen_enformer_full():
  """The headline network (decode.py:78-80: EnformerTrunk(7, 1536, 11, 8, 64) +
  ConvHead(1, 3072)) and the DNA reward oracle's 3-task head (oracle.py:72) on SVDD-step-shaped
  candidate sets: 32 sequences x M = 10 candidates, L = 200 (helpers.svdd_step_candidates)."""
  d = ref_diffusion(200)
  S, M, L = 32, 10, 200
  x, cand = helpers.svdd_step_candidates(S, M, L, seed=2024)
  emb, head = ref_enformer(full=True, calibrate=True)
  orc = ref_dna_reward_model(emb)
  vals, vals3 = [], []
  with torch.no_grad():
    for m in range(M):                           # the reference scores candidate by candidate (:1203-1210)
      oh = d.transform_samples(cand[m]).float()
      vals.append(head(emb(oh)).squeeze())                   # diffusion_gosai.py:1208-1209
    for m in range(2):                           # reward oracle: reward_model(x.transpose(1, 2)) -> [N, 3, 1]; [:, 0] is scored (:1430)
      oh = d.transform_samples(cand[m]).float()
      vals3.append(orc(oh.transpose(1, 2)).squeeze(-1))
  save('enformer_full.npz', x=x.to(torch.int8), cand=cand.to(torch.int8), values=torch.stack(vals),
       values3=torch.stack(vals3), checksum=helpers.state_checksum(emb.state_dict()),
       head3_checksum=helpers.state_checksum(orc.head.state_dict()))


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


def gen_dna_trajectories():
  """BASELINE configs 2 and 3 in small: the reference's own controlled_sample (SVDD-MC, full
  Enformer value net) and controlled_sample_tweedie (SVDD-PM, 3-task Enformer reward oracle,
  task 0 scored) at L = 200, M = 10, plus one controlled step of each kind from a half-unmasked
  state.  The uniform tensors are INJECTED from a seeded generator (torch.rand_like patched), so
  only seeds and tokens are stored."""
  out = {}
  d = ref_diffusion(200)
  emb, head = ref_enformer(full=True, calibrate=False)
  orc = ref_dna_reward_model(emb)
  L, M = 200, 10
  for tag, B, steps, seed in (('mc', 3, 4, 901), ('pm', 2, 3, 902)):
    U = torch.rand(steps, M, B, L, 5, generator=torch.Generator().manual_seed(seed))
    with RandLikeTap(inject=list(U.reshape(steps * M, B, L, 5))), torch.no_grad():
      if tag == 'mc':
        x = d.controlled_sample(emb, head, num_steps=steps, eval_sp_size=B, sample_M=M)
      else:
        x = d.controlled_sample_tweedie(orc, num_steps=steps, eval_sp_size=B, sample_M=M,
                                        options='True', task='dna')
    out[f'{tag}_tokens'], out[f'{tag}_seed'], out[f'{tag}_shape'] = x, seed, np.int64([steps, M, B, L])
  # single controlled steps through the reference's step functions (the drop-in surface, A15)
  timesteps = torch.linspace(1, 1e-5, 129)
  dt = (1 - 1e-5) / 128
  B, step = 3, 70
  x = helpers.random_tokens(B, L, 3070, 0.5)
  U = torch.rand(M, B, L, 5, generator=torch.Generator().manual_seed(903))
  t = timesteps[step] * torch.ones(B, 1)
  with torch.no_grad():
    with RandLikeTap(inject=list(U)):
      a, _, qa, ca = d._ddpm_update_finetune_controlled(x, t, dt, emb, head, repeats=M)
    with RandLikeTap(inject=list(U)):
      b, _, qb, cb = d._ddpm_update_finetune_controlled_twedie(x, t, dt, orc, repeats=M, options='True', task='dna')
    with RandLikeTap(inject=list(U)):
      c, _, _, _ = d._ddpm_update_finetune_controlled_twedie(x, t, dt, orc, repeats=M, options='False', task='dna')
  assert torch.equal(qa, qb)
  out.update(step_x=x, step_seed=903, step_index=step, step_mc_next=a, step_pm_next=b, step_pm_raw_next=c,
             step_q=qa, step_copy_flag=ca)
  save('dna_trajectories.npz', **out)


def ref_dit(n_blocks, length=200, time_conditioning=False):
  """The reference's Diffusion with ``backbone: dit`` cannot be constructed (models/__init__.py
  comments ``dit`` out, diffusion_gosai.py:102-104 would raise AttributeError), so the golden
  builds the reference's own ``models.dit.DIT`` with the same seed the svdd_b200 container uses
  (helpers.build_dit: torch.manual_seed(44) then Diffusion(cfg) -> the backbone is the first
  module constructed) and drives its sub-modules in the order of DIT.forward's body, whose
  ``return x`` is missing in the reference (models/dit.py:355-366)."""
  dit = ref_import.import_reference_dit()
  NS = __import__('types').SimpleNamespace
  cfg = NS(model=NS(hidden_size=768, cond_dim=128, n_blocks=n_blocks, n_heads=12, dropout=0.1,
                    scale_by_sigma=True, length=length))
  torch.manual_seed(44)
  m = dit.DIT(cfg, vocab_size=5).eval()
  helpers.perturb_dit_(m, 13)

  def forward(indices, sigma):
    x = m.vocab_embed(indices)
    c = torch.nn.functional.silu(m.sigma_map(sigma))
    rotary_cos_sin = m.rotary_emb(x)
    for blk in m.blocks:
      x = blk(x, rotary_cos_sin, c, seqlens=None)
    return m.output_layer(x, c)
  return m, forward


def gen_dit():
  out = {}
  for tag, n_blocks, L, B, sigma in (('b2_L200', 2, 200, 3, 0.0), ('b12_L200', 12, 200, 2, 0.0),
                                     ('b2_L50_sigma', 2, 50, 3, 0.7), ('b1_L333', 1, 333, 1, 0.0)):
    m, fwd = ref_dit(n_blocks, L)
    x = helpers.random_tokens(B, L, 500 + L + n_blocks, 0.5)
    x[0] = 4
    with torch.no_grad():
      logits = fwd(x, torch.full((B,), sigma))
    out[f'{tag}_tokens'], out[f'{tag}_logits'], out[f'{tag}_sigma'] = x, logits, np.float32(sigma)
    out[f'{tag}_checksum'] = helpers.state_checksum(m.state_dict())
  save('dit_seed44.npz', **out)


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
  gen_enformer_full()
  gen_trajectories()
  gen_dna_trajectories()
  gen_dit()
  gen_ddpm_cache()
