"""-m gpu: the whole sampler through the reference-shaped API
(svdd_b200.diffusion_gosai.Diffusion) against the oracle / reference goldens."""
import numpy as np
import pytest
import torch

import helpers
from oracle import nets, svdd
from svdd_b200 import _lib, config, diffusion_gosai, value_nets

pytestmark = pytest.mark.gpu


def T(a):
  return torch.from_numpy(np.asarray(a))


class ReplayDenoiser:
  """Test double for the packed denoiser: replays recorded logits call by call."""

  def __init__(self, logits_seq):
    self.seq, self.i = logits_seq, 0

  def forward(self, tokens, sigma=0.0, out=None):
    lg = self.seq[self.i]
    self.i += 1
    assert lg.shape[0] == tokens.shape[0]
    if out is not None:
      out.copy_(lg)
      return out
    return lg.clone()


class ReplayScorer:
  def __init__(self, scores_seq):
    self.seq, self.i = scores_seq, 0

  def score(self, tokens, out=None):
    s = self.seq[self.i].reshape(-1)
    self.i += 1
    if out is not None:
      out.copy_(s)
      return out
    return s.clone()


def _rna_model(cuda):
  cfg = config.load_config('rna')
  torch.manual_seed(44)
  return diffusion_gosai.Diffusion(cfg).to(cuda).eval()


def _oracle_fns():
  den = helpers.build_denoiser(44, 50)
  sd = {'backbone.' + k: v for k, v in den.state_dict().items()}
  emb, head = helpers.build_convgru_value()
  oe, oh = helpers.build_convgru_oracle()
  denoiser = lambda x: nets.denoiser_logits(sd, x)
  value = lambda tok: nets.convgru_value(emb.state_dict(), head.state_dict(),
                                         svdd.transform_samples(tok).float()).squeeze()
  reward = lambda tok: nets.convgru_value(oe.state_dict(), oh.state_dict(),
                                          svdd.transform_samples(tok).float()).squeeze()
  return denoiser, value, reward


def test_backbone_weights_match_reference_seed(cuda):
  m = _rna_model(cuda)
  g = helpers.load_golden('denoiser_seed44.npz')
  np.testing.assert_array_equal(helpers.state_checksum({k: v.cpu() for k, v in m.backbone.state_dict().items()}),
                                g['L50_checksum'])


def test_mc_trajectory_bit_exact_with_reference_logits_values_noise(cuda):
  """north_star: "output token sequences are bit-exact when fed the reference's
  logits, values and uniform-noise tensors" -- the golden trajectory was produced
  by the reference's own controlled_sample; the oracle trace supplies its
  per-step logits and values; the CUDA engine must land on the same tokens."""
  g = helpers.load_golden('trajectories.npz')
  denoiser, value, _ = _oracle_fns()
  trace = []
  with torch.no_grad():
    x_ref = svdd.controlled_sample(denoiser, value, B=4, L=50, M=3, num_steps=12,
                                   noise=svdd.ArrayNoise(g['mc_U']), trace=trace)
    final_logits = denoiser(trace[-1]['x_next'])
  assert np.array_equal(x_ref.numpy(), g['mc_tokens'])
  m = _rna_model(cuda)
  fake_den = ReplayDenoiser([t['logits'].to(cuda) for t in trace] + [final_logits.to(cuda)])
  fake_val = ReplayScorer([t['scores'].t().contiguous().to(cuda) for t in trace])
  m.backbone.packed = lambda: fake_den
  x = m.controlled_sample(fake_val, None, num_steps=12, eval_sp_size=4, sample_M=3,
                          noise=diffusion_gosai.InjectedNoise(T(g['mc_U']).to(cuda)))
  np.testing.assert_array_equal(x.cpu().numpy(), g['mc_tokens'])


def test_pm_trajectory_bit_exact_with_reference_logits_values_noise(cuda):
  g = helpers.load_golden('trajectories.npz')
  denoiser, _, reward = _oracle_fns()
  # record, per step, the logits of x and of every candidate plus the rewards
  calls = []
  rec_den = lambda x: (calls.append(None), denoiser(x))[1]
  trace = []
  with torch.no_grad():
    x_ref = svdd.controlled_sample_tweedie(denoiser, reward, B=4, L=50, M=3, num_steps=6,
                                           noise=svdd.ArrayNoise(g['pm_U']), trace=trace)
    assert np.array_equal(x_ref.numpy(), g['pm_tokens'])
    den_seq = []
    for t in trace:
      den_seq.append(t['logits'].to(cuda))
      den_seq.append(torch.cat([denoiser(t['cand'][m]) for m in range(3)], 0).to(cuda))
    den_seq.append(denoiser(trace[-1]['x_next']).to(cuda))
  m = _rna_model(cuda)
  fake_den = ReplayDenoiser(den_seq)
  fake_rew = ReplayScorer([t['scores'].t().contiguous().to(cuda) for t in trace])
  m.backbone.packed = lambda: fake_den
  x = m.controlled_sample_tweedie(fake_rew, num_steps=6, eval_sp_size=4, sample_M=3, options='True',
                                  task='rna', noise=diffusion_gosai.InjectedNoise(T(g['pm_U']).to(cuda)))
  np.testing.assert_array_equal(x.cpu().numpy(), g['pm_tokens'])


def test_plain_and_sample_with_reference_logits(cuda):
  g = helpers.load_golden('trajectories.npz')
  denoiser, _, _ = _oracle_fns()
  # replay the oracle's logits along the golden plain trajectory
  sched, _ = svdd.move_chances(16, 1e-5)
  x = torch.full((4, 50), 4, dtype=torch.int64)
  seq = []
  with torch.no_grad():
    for i in range(16):
      lg = denoiser(x)
      seq.append(lg.to(cuda))
      q = svdd.build_q_xs(svdd.subs_parameterization(lg, x), sched[i, 0], sched[i, 1])
      x = svdd.draw_candidates(x, q, T(g['plain_U'][i]))[0]
    seq.append(denoiser(x).to(cuda))
  m = _rna_model(cuda)
  fake = ReplayDenoiser(seq)
  m.backbone.packed = lambda: fake
  out = m.decode_sample(num_steps=16, eval_sp_size=4,
                        noise=diffusion_gosai.InjectedNoise(T(g['plain_U']).to(cuda)))
  np.testing.assert_array_equal(out.cpu().numpy(), g['plain_tokens'])


def test_ddpm_cache_predictor(cuda):
  """sampling.predictor == 'ddpm_cache' (diffusion_gosai.py:755-773, 858-865) in _sample:
  (1) replaying the reference's fp32 logits call by call reproduces the golden tokens, every
  intermediate state and the number of denoiser forwards (78 of 161: the cache skips stage 1
  while the batch is unchanged); (2) with the tensor-core denoiser in the loop every
  transition is the oracle's given the engine's own log-probabilities."""
  g = helpers.load_golden('ddpm_cache.npz')
  denoiser, _, _ = _oracle_fns()
  U = T(g['U'])
  seq = []

  def recording_log_p(x):
    lg = denoiser(x)
    seq.append(lg.to(cuda))
    return svdd.subs_parameterization(lg, x)
  with torch.no_grad():
    x_o, mid_o, n_o = svdd.sample_ddpm_cache(None, 2, 50, 160, noise=svdd.ArrayNoise(g['U']),
                                             log_p_fn=recording_log_p)
  np.testing.assert_array_equal(x_o.numpy(), g['tokens'])
  m = _rna_model(cuda)
  m.sampler = 'ddpm_cache'
  fake = ReplayDenoiser(seq)
  m.backbone.packed = lambda: fake
  out, mids = m._sample(num_steps=160, eval_sp_size=2, noise=diffusion_gosai.InjectedNoise(U.to(cuda)))
  np.testing.assert_array_equal(out.cpu().numpy(), g['tokens'])
  np.testing.assert_array_equal(torch.stack(mids).cpu().numpy(), g['mid'])
  assert m.last_denoiser_forwards == int(g['n_forward']) == len(seq)
  fake.i = 0                                            # decode_sample takes the same branch (:912-919)
  out_d = m.decode_sample(num_steps=160, eval_sp_size=2, noise=diffusion_gosai.InjectedNoise(U.to(cuda)))
  np.testing.assert_array_equal(out_d.cpu().numpy(), g['tokens'])

  m2 = _rna_model(cuda)
  m2.sampler = 'ddpm_cache'
  out2, mids2 = m2._sample(num_steps=160, eval_sp_size=2, noise=diffusion_gosai.InjectedNoise(U.to(cuda)))
  sig = torch.zeros(2, device=cuda)
  with torch.no_grad():
    x_e, mid_e, n_e = svdd.sample_ddpm_cache(None, 2, 50, 160, noise=svdd.ArrayNoise(g['U']),
                                             log_p_fn=lambda x: m2.forward(x.to(cuda), sig).cpu())
  np.testing.assert_array_equal(out2.cpu().numpy(), x_e.numpy())
  np.testing.assert_array_equal(torch.stack(mids2).cpu().numpy(), torch.stack(mid_e).numpy())
  assert m2.last_denoiser_forwards == n_e < 161


@pytest.mark.parametrize('mode', ['mc', 'pm', 'pm_raw'])
def test_real_networks_every_transition_matches_oracle(cuda, mode):
  """With the tensor-core networks in the loop, every reverse-step transition is the
  oracle's given the kernels' own logits / values and the injected noise; the
  networks themselves stay within the stated tolerance of the fp32 oracle."""
  m = _rna_model(cuda)
  B, M, steps = 6, 5, 10
  U = torch.rand(steps, M, B, 50, 5, generator=torch.Generator().manual_seed(5))
  noise = diffusion_gosai.InjectedNoise(U.to(cuda))
  trace = []
  if mode == 'mc':
    emb, head = helpers.build_convgru_value()
    x = m.controlled_sample(emb.to(cuda), head.to(cuda), num_steps=steps, eval_sp_size=B,
                            sample_M=M, noise=noise, trace=trace)
  else:
    emb, head = helpers.build_convgru_oracle()
    rm = value_nets.OriBaseModel(emb.to(cuda), head.to(cuda))
    x = m.controlled_sample_tweedie(rm, num_steps=steps, eval_sp_size=B, sample_M=M, noise=noise,
                                    options='True' if mode == 'pm' else 'False', task='rna', trace=trace)
  assert x.shape == (B, 50) and int(x.max()) <= 3
  denoiser, value, reward = _oracle_fns()
  for rec in trace[:-1]:
    xs, lg = rec['x'].cpu().long(), rec['logits'].cpu()
    q = svdd.build_q_xs(svdd.subs_parameterization(lg, xs), rec['mc_t'], rec['mc_s'])
    cand = svdd.draw_candidates(xs, q, U[rec['step']])
    assert torch.equal(cand, rec['cand'].cpu().long())
    if mode == 'pm':
      x0 = svdd.subs_parameterization(rec['logits2'].cpu(), cand.reshape(M * B, 50)).argmax(-1)
      assert torch.equal(x0, rec['x0'].cpu().long())
    idx = svdd.select(rec['scores'].cpu().t().contiguous())
    assert torch.equal(svdd.gather_selected(cand, idx), rec['x_next'].cpu().long())
    with torch.no_grad():
      ref = denoiser(xs)
    assert float((lg - ref).abs().max() / ref.abs().max()) < 3e-2
  rec = trace[3]
  toks = (rec['x0'] if mode == 'pm' else rec['cand']).cpu().long().reshape(M * B, 50)
  with torch.no_grad():
    ref = (value if mode == 'mc' else reward)(toks).reshape(M, B)
  assert float((rec['scores'].cpu() - ref).abs().max()) < 1e-2


def test_graph_replay_equals_eager_and_reseeds(cuda):
  m = _rna_model(cuda)
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  m.manual_seed(1234)
  m.use_cuda_graph = False
  a = m.controlled_sample(emb, head, num_steps=8, eval_sp_size=16, sample_M=4)
  b = m.controlled_sample(emb, head, num_steps=8, eval_sp_size=16, sample_M=4)
  m.manual_seed(1234)
  m.use_cuda_graph = True
  c = m.controlled_sample(emb, head, num_steps=8, eval_sp_size=16, sample_M=4)   # capture + replay
  d = m.controlled_sample(emb, head, num_steps=8, eval_sp_size=16, sample_M=4)   # replay, new key
  assert torch.equal(a, c) and torch.equal(b, d) and not torch.equal(a, b)


def test_philox_run_matches_oracle_philox_noise(cuda):
  """Product path (in-kernel noise): the oracle fed the same counter-based stream
  reproduces every transition."""
  m = _rna_model(cuda)
  emb, head = helpers.build_convgru_value()
  m.manual_seed(99)
  m.use_cuda_graph = False
  trace = []
  B, M, steps = 5, 4, 6
  m.controlled_sample(emb.to(cuda), head.to(cuda), num_steps=steps, eval_sp_size=B, sample_M=M,
                      row_offset=7, trace=trace)
  noise = svdd.PhiloxNoise(99, row_offset=7)
  for rec in trace[:-1]:
    xs, lg = rec['x'].cpu().long(), rec['logits'].cpu()
    q = svdd.build_q_xs(svdd.subs_parameterization(lg, xs), rec['mc_t'], rec['mc_s'])
    cand = svdd.draw_candidates(xs, q, noise.draws(rec['step'], M, B, 50))
    assert torch.equal(cand, rec['cand'].cpu().long())


def test_alpha_soft_selection(cuda):
  m = _rna_model(cuda)
  emb, head = helpers.build_convgru_value()
  B, M, steps = 6, 8, 5
  g = torch.Generator().manual_seed(8)
  U = torch.rand(steps, M, B, 50, 5, generator=g)
  Us = torch.rand(steps, B, M, generator=g)
  for alpha in (0.1, 1.0):
    trace = []
    m.controlled_sample(emb.to(cuda), head.to(cuda), num_steps=steps, eval_sp_size=B, sample_M=M,
                        alpha=alpha, noise=diffusion_gosai.InjectedNoise(U.to(cuda), Us.to(cuda)),
                        trace=trace)
    for rec in trace[:-1]:
      idx = svdd.select(rec['scores'].cpu().t().contiguous(), alpha, Us[rec['step']])
      assert torch.equal(svdd.gather_selected(rec['cand'].cpu().long(), idx), rec['x_next'].cpu().long())


def test_forward_api_and_step_functions(cuda):
  m = _rna_model(cuda)
  x = helpers.random_tokens(3, 50, 2, 0.5).to(cuda)
  lp = m.forward(x, torch.zeros(3, device=cuda))
  assert lp.shape == (3, 50, 5)
  keep = (x != 4).cpu()
  p = lp.exp().cpu()
  assert torch.equal(p[keep].argmax(-1), x.cpu()[keep]) and float(p[keep].sum(-1).max()) == 1.0
  t = torch.full((3, 1), 0.5, device=cuda)
  xn, x_in, q, copy_flag = m._ddpm_update_finetune(x, t, (1 - 1e-5) / 128)
  assert torch.equal(xn.cpu()[keep], x.cpu()[keep]) and q.shape == (3, 50, 5)
  assert torch.equal(copy_flag.cpu(), keep.long())


def test_smoke_entry(cuda):
  import __graft_entry__
  __graft_entry__._smoke_on(cuda)


def test_cli_decode_and_tweedie_write_reference_npz_layout(cuda, tmp_path):
  """decode.py / decode_tweedie.py end to end (random-init RNA nets, small batch):
  outputs ./log/{task}-{reward_name}[_tw].npz with float32 (N,) arrays."""
  import subprocess, sys
  for script, extra, name in (('decode.py', [], 'rna-MRL.npz'),
                              ('decode_tweedie.py', ['--tweedie', 'True'], 'rna-MRL_tw.npz')):
    cmd = [sys.executable, script, '--task', 'rna', '--sample_M', '3', '--batch_size', '8',
           '--val_batch_num', '1', '--reward_name', 'MRL', '--random_init', '--out_dir', str(tmp_path)] + extra
    r = subprocess.run(cmd, cwd=helpers.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    d = np.load(tmp_path / name)
    assert sorted(d.files) == ['baseline', 'decoding']
    assert d['decoding'].shape == (8,) and d['decoding'].dtype == np.float32
    assert d['baseline'].shape == (8,) and np.isfinite(d['decoding']).all()
    assert 'timing: SVDD 1 x 8 sequences' in r.stdout and 'baseline 3 x 8 rollouts' in r.stdout


def test_base_model_tuple_and_svdd_gain(cuda):
  """controlled_decode returns the reference's 5-tuple; with a value net equal to the
  reward oracle, SVDD-MC's rewards beat the unguided baseline (sanity of the whole loop)."""
  from svdd_b200.base_model import BaseModel
  torch.manual_seed(44)
  m = BaseModel(None, None, cdq=False, batch_size=32, val_batch_num=1, task='rna', random_init=True)
  m.embedding, m.head = m.reward_model.embedding, m.reward_model.head   # value == reward
  m = m.to(cuda).eval()
  samples, v, r, top, base = m.controlled_decode(gen_batch_num=1, sample_M=8)
  assert samples[0].shape == (32, 50) and v.shape == r.shape == base.shape == (32,)
  assert top.shape == (32,) and torch.allclose(v, r)
  assert float(r.mean()) > float(base.mean())
  assert m.timing['svdd_batches'] == 1 and m.timing['baseline_rollouts'] == 8
  assert m.timing['svdd_s'] > 0 and m.timing['baseline_s'] > 0


# ---------------------------------------------------------------------------------------------
# DNA (L = 200, M = 10): BASELINE configs 2 (SVDD-MC, Enformer value net) and 3 (SVDD-PM, 3-task
# Enformer reward oracle).  Goldens: tests/golden/dna_trajectories.npz (reference's own code).
# ---------------------------------------------------------------------------------------------

def _dna_model(cuda):
  cfg = config.load_config('dna')
  torch.manual_seed(44)
  return diffusion_gosai.Diffusion(cfg).to(cuda).eval()


def _dna_oracle_fns():
  from test_oracle_golden import dna_fns
  return dna_fns()


def _dna_noise(g, tag):
  from test_oracle_golden import dna_noise
  return dna_noise(g, tag)


def test_dna_mc_trajectory_bit_exact_with_reference_logits_values_noise(cuda):
  """c2-shaped: the reference's controlled_sample with the FULL Enformer value net produced the
  golden tokens; the oracle trace supplies its per-step logits / values; the CUDA engine lands
  on the same tokens."""
  g = helpers.load_golden('dna_trajectories.npz')
  denoiser, value, _ = _dna_oracle_fns()
  U, (steps, M, B, L) = _dna_noise(g, 'mc')
  trace = []
  with torch.no_grad():
    x_ref = svdd.controlled_sample(denoiser, value, B=B, L=L, M=M, num_steps=steps,
                                   noise=svdd.ArrayNoise(U), trace=trace)
    final_logits = denoiser(trace[-1]['x_next'])
  assert np.array_equal(x_ref.numpy(), g['mc_tokens'])
  m = _dna_model(cuda)
  fake_den = ReplayDenoiser([t['logits'].to(cuda) for t in trace] + [final_logits.to(cuda)])
  fake_val = ReplayScorer([t['scores'].t().contiguous().to(cuda) for t in trace])
  m.backbone.packed = lambda: fake_den
  x = m.controlled_sample(fake_val, None, num_steps=steps, eval_sp_size=B, sample_M=M,
                          noise=diffusion_gosai.InjectedNoise(U.to(cuda)))
  np.testing.assert_array_equal(x.cpu().numpy(), g['mc_tokens'])


def test_dna_pm_trajectory_bit_exact_with_reference_logits_values_noise(cuda):
  """c3-shaped: controlled_sample_tweedie with the 3-task reward oracle (task 0 scored)."""
  g = helpers.load_golden('dna_trajectories.npz')
  denoiser, _, reward = _dna_oracle_fns()
  U, (steps, M, B, L) = _dna_noise(g, 'pm')
  trace = []
  with torch.no_grad():
    x_ref = svdd.controlled_sample_tweedie(denoiser, reward, B=B, L=L, M=M, num_steps=steps,
                                           noise=svdd.ArrayNoise(U), trace=trace)
    assert np.array_equal(x_ref.numpy(), g['pm_tokens'])
    den_seq = []
    for t in trace:
      den_seq.append(t['logits'].to(cuda))
      den_seq.append(torch.cat([denoiser(t['cand'][k]) for k in range(M)], 0).to(cuda))
    den_seq.append(denoiser(trace[-1]['x_next']).to(cuda))
  m = _dna_model(cuda)
  fake_den = ReplayDenoiser(den_seq)
  fake_rew = ReplayScorer([t['scores'].t().contiguous().to(cuda) for t in trace])
  m.backbone.packed = lambda: fake_den
  x = m.controlled_sample_tweedie(fake_rew, num_steps=steps, eval_sp_size=B, sample_M=M, options='True',
                                  task='dna', noise=diffusion_gosai.InjectedNoise(U.to(cuda)))
  np.testing.assert_array_equal(x.cpu().numpy(), g['pm_tokens'])


@pytest.mark.parametrize('mode', ['mc', 'pm'])
def test_dna_real_networks_every_transition_matches_oracle(cuda, mode):
  """c2 / c3 shaped trace (B = 8, M = 10, L = 200, 4 steps, injected noise) with the tensor-core
  networks in the loop -- the fused denoiser and the FULL Enformer value net (mc) / 3-task
  Enformer reward oracle (pm): every transition is the oracle's given the kernels' own logits /
  values; x0 equals the oracle's Tweedie argmax; the scores are task 0's and sit within the
  stated tolerance of the fp32 oracle."""
  from svdd_b200 import synthetic
  m = _dna_model(cuda)
  B, M, L, steps = 8, 10, 200, 4
  U = torch.rand(steps, M, B, L, 5, generator=torch.Generator().manual_seed(15))
  noise = diffusion_gosai.InjectedNoise(U.to(cuda))
  trace = []
  if mode == 'mc':
    emb, head = helpers.build_enformer(full=True)
    x = m.controlled_sample(emb.to(cuda), head.to(cuda), num_steps=steps, eval_sp_size=B,
                            sample_M=M, noise=noise, trace=trace)
  else:
    rm = synthetic.build_dna_reward_model().to(cuda)
    x = m.controlled_sample_tweedie(rm, num_steps=steps, eval_sp_size=B, sample_M=M, noise=noise,
                                    options='True', task='dna', trace=trace)
  assert x.shape == (B, L) and int(x.max()) <= 3
  denoiser, value, reward = _dna_oracle_fns()
  for rec in trace[:-1]:
    xs, lg = rec['x'].cpu().long(), rec['logits'].cpu()
    q = svdd.build_q_xs(svdd.subs_parameterization(lg, xs), rec['mc_t'], rec['mc_s'])
    cand = svdd.draw_candidates(xs, q, U[rec['step']])
    assert torch.equal(cand, rec['cand'].cpu().long())
    if mode == 'pm':
      x0 = svdd.subs_parameterization(rec['logits2'].cpu(), cand.reshape(M * B, L)).argmax(-1)
      assert torch.equal(x0, rec['x0'].cpu().long())
    idx = svdd.select(rec['scores'].cpu().t().contiguous())
    assert torch.equal(svdd.gather_selected(cand, idx), rec['x_next'].cpu().long())
    with torch.no_grad():
      ref = denoiser(xs)
    assert float((lg - ref).abs().max() / ref.abs().max()) < 3e-2
  last = trace[-1]
  assert torch.equal(svdd.subs_parameterization(last['logits'].cpu(), last['x'].cpu().long())[:, :, :4].argmax(-1), x.cpu())
  rec = trace[2]
  toks = (rec['x0'] if mode == 'pm' else rec['cand']).cpu().long().reshape(M * B, L)
  with torch.no_grad():
    ref = (value if mode == 'mc' else reward)(toks).reshape(M, B)
  err = float((rec['scores'].cpu() - ref).abs().max())
  agree = float((rec['scores'].cpu().argmax(0) == ref.argmax(0)).float().mean())
  print(f'\n[dna {mode}] step-2 scores: max|d| vs fp32 oracle {err:.3e}, argmax agreement {agree:.2f}')
  assert err < 3.5e-2


def test_controlled_step_functions_match_reference_known_answers(cuda):
  """A15: _ddpm_update_finetune_controlled / _controlled_twedie -- the reference's step
  functions, part of the drop-in surface -- return the reference's (final_samples, x, q_xs,
  copy_flag) on the golden stage cases a-d (tests/golden/stage_kats.npz: injected logits, noise
  and scores through the reference's own function)."""
  g = helpers.load_golden('stage_kats.npz')
  timesteps = torch.linspace(1, 1e-5, 129)
  dt = (1 - 1e-5) / 128
  for tag in 'abcd':
    c = {k[2:]: g[k] for k in g.files if k.startswith(tag + '_')}
    x, logits, U, scores = T(c['x']), T(c['logits']), T(c['U']), T(c['scores'])
    B, L = x.shape
    M = U.shape[0]
    cfg = config.load_config('rna')
    cfg.model.length = L
    model = diffusion_gosai.Diffusion(cfg).to(cuda).eval()
    t = (timesteps[int(c['step'])] * torch.ones(B, 1)).to(cuda)
    fake = ReplayDenoiser([logits.to(cuda)])
    model.backbone.packed = lambda: fake
    scorer = ReplayScorer([scores.t().contiguous().to(cuda)])
    final, x_in, q, copy_flag = model._ddpm_update_finetune_controlled(
        x.to(cuda), t, dt, scorer, None, repeats=M, U=U.to(cuda))
    np.testing.assert_array_equal(final.cpu().numpy(), c['x_next'], err_msg=tag)
    np.testing.assert_array_equal(x_in.cpu().numpy(), c['x'], err_msg=tag)
    np.testing.assert_allclose(q.cpu().numpy(), c['q'], rtol=2e-6, atol=0, err_msg=tag)
    np.testing.assert_array_equal(copy_flag.cpu().numpy(), (c['x'] != 4).astype(np.int64), err_msg=tag)
    assert final.dtype == torch.int64 and copy_flag.dtype == torch.int64
    # the Tweedie variant with options='False' scores the raw candidates: same golden answer
    fake.i, scorer.i = 0, 0
    rm = value_nets.OriBaseModel(scorer, None)
    final2, _, q2, _ = model._ddpm_update_finetune_controlled_twedie(
        x.to(cuda), t, dt, rm, repeats=M, options='False', task='rna', U=U.to(cuda))
    np.testing.assert_array_equal(final2.cpu().numpy(), c['x_next'], err_msg=tag)
    assert torch.equal(q2, q)


def test_controlled_step_functions_dna_golden(cuda):
  """The same three step functions at L = 200 with the reference's REAL networks behind the
  golden (full Enformer value net / 3-task reward oracle, options 'True' and 'False'): replaying
  the oracle's logits and scores through the CUDA step functions reproduces the reference's
  next state."""
  g = helpers.load_golden('dna_trajectories.npz')
  denoiser, value, reward = _dna_oracle_fns()
  sched, _ = svdd.move_chances(128, 1e-5)
  i = int(g['step_index'])
  xs = T(g['step_x'])
  B, L, M = 3, 200, 10
  U = torch.rand(M, B, L, 5, generator=torch.Generator().manual_seed(int(g['step_seed'])))
  t = (torch.linspace(1, 1e-5, 129)[i] * torch.ones(B, 1)).to(cuda)
  dt = (1 - 1e-5) / 128
  m = _dna_model(cuda)
  with torch.no_grad():
    lg = denoiser(xs)
    cand = svdd.draw_candidates(xs, svdd.build_q_xs(svdd.subs_parameterization(lg, xs), float(sched[i, 0]),
                                                    float(sched[i, 1])), U)
    lg2 = torch.cat([denoiser(cand[k]) for k in range(M)], 0)
    x0 = torch.stack([svdd.tweedie_onehot(denoiser, cand[k]) for k in range(M)])
    s_mc = torch.stack([value(cand[k]) for k in range(M)])
    s_pm = torch.stack([reward(x0[k]) for k in range(M)])
    s_raw = torch.stack([reward(cand[k]) for k in range(M)])
  for fn, den_seq, sc, key in (('mc', [lg], s_mc, 'step_mc_next'), ('True', [lg, lg2], s_pm, 'step_pm_next'),
                               ('False', [lg], s_raw, 'step_pm_raw_next')):
    fake = ReplayDenoiser([d.to(cuda) for d in den_seq])
    m.backbone.packed = lambda fake=fake: fake
    scorer = ReplayScorer([sc.to(cuda)])
    if fn == 'mc':
      out = m._ddpm_update_finetune_controlled(xs.to(cuda), t, dt, scorer, None, repeats=M, U=U.to(cuda))
    else:
      out = m._ddpm_update_finetune_controlled_twedie(xs.to(cuda), t, dt, value_nets.OriBaseModel(scorer, None),
                                                      repeats=M, options=fn, task='dna', U=U.to(cuda))
    np.testing.assert_array_equal(out[0].cpu().numpy(), g[key], err_msg=fn)
    np.testing.assert_allclose(out[2].cpu().numpy(), g['step_q'], rtol=2e-6, atol=0)
    np.testing.assert_array_equal(out[3].cpu().numpy(), g['step_copy_flag'])


@pytest.mark.parametrize('mode,alpha', [('mc', 0.0), ('mc', 0.1), ('pm', 0.0), ('pm', 0.1), ('plain', 0.0)])
@pytest.mark.parametrize('graph', [False, True])
def test_shard_equivalence_on_one_gpu(cuda, mode, alpha, graph):
  """results do not depend on how the batch is sharded.  concat(run(rows 0..3),
  run(rows 4..7, row_offset = 4)) == run(rows 0..7) for SVDD-MC, SVDD-PM (incl. alpha > 0, whose
  selection noise is keyed by the global row too) and plain sampling, with the in-kernel Philox
  stream, eager and through the CUDA graph (row offset in device memory: ONE captured graph
  serves both halves)."""
  m = _rna_model(cuda)
  m.use_cuda_graph = graph
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  oe, oh = helpers.build_convgru_oracle()
  rm = value_nets.OriBaseModel(oe.to(cuda), oh.to(cuda))
  steps, M = 6, 4

  def run(rows, off):
    m.manual_seed(77)                          # the same run key for every shard
    if mode == 'mc':
      return m.controlled_sample(emb, head, num_steps=steps, eval_sp_size=rows, sample_M=M, alpha=alpha,
                                 row_offset=off)
    if mode == 'pm':
      return m.controlled_sample_tweedie(rm, num_steps=steps, eval_sp_size=rows, sample_M=M, alpha=alpha,
                                         options='True', task='rna', row_offset=off)
    return m.decode_sample(num_steps=steps, eval_sp_size=rows, row_offset=off)

  whole = run(8, 0)
  captured = len(m._graphs)
  halves = torch.cat([run(4, 0), run(4, 4)], 0)
  assert torch.equal(whole, halves)
  uneven = torch.cat([run(3, 0), run(5, 3)], 0)
  assert torch.equal(whole, uneven)
  assert not torch.equal(run(4, 0), run(4, 4))
  if graph:
    assert len(m._graphs) == captured + 3, 'one graph per batch shape, none per row offset'


def test_graph_cache_survives_workspace_growth_and_handle_rebuild(cuda):
  """a cached graph points into the handles' workspaces and packed weights.  Growing
  a workspace (a larger scoring call between two decodes) or re-packing the value net must not
  invalidate a graph that is replayed afterwards."""
  m = _rna_model(cuda)
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  m.manual_seed(5)
  a = m.controlled_sample(emb, head, num_steps=6, eval_sp_size=8, sample_M=4)
  scorer = value_nets.packed_scorer(emb, head)
  ws_before = scorer._ws
  big = helpers.random_tokens(40000, 50, 1, 0.5).to(cuda)
  value_nets.score_tokens(emb, head, big)                       # grows the scorer's workspace
  m.backbone.packed().forward(big[:20000], 0.0)                 # and the denoiser's
  assert scorer._ws is not ws_before
  junk = [torch.full((1 << 20,), 7, dtype=torch.uint8, device=cuda) for _ in range(64)]   # reuse freed blocks
  m.manual_seed(5)
  b = m.controlled_sample(emb, head, num_steps=6, eval_sp_size=8, sample_M=4)   # replay of the cached graph
  assert torch.equal(a, b)
  del junk
  # re-packing (weights changed) makes a new handle with a new uid -> a new graph, the old one stays valid
  with torch.no_grad():
    head.channel_transform.conv.layer.bias.add_(1.0)
  m.manual_seed(5)
  c = m.controlled_sample(emb, head, num_steps=6, eval_sp_size=8, sample_M=4)
  assert value_nets.packed_scorer(emb, head).uid != scorer.uid and len(m._graphs) == 2
  assert torch.equal(a, c)                                      # a constant shift does not change the argmax


def test_time_conditioning_graph_capture(cuda):
  """with time_conditioning=True every step has its own sigma; the time-bias rows
  are computed before the capture (a miss inside it raises) and pinned with the graph."""
  cfg = config.load_config('rna')
  cfg.time_conditioning = True
  torch.manual_seed(44)
  m = diffusion_gosai.Diffusion(cfg).to(cuda).eval()
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  m.manual_seed(3)
  m.use_cuda_graph = False
  a = m.controlled_sample(emb, head, num_steps=9, eval_sp_size=5, sample_M=3)
  m.manual_seed(3)
  m.use_cuda_graph = True
  b = m.controlled_sample(emb, head, num_steps=9, eval_sp_size=5, sample_M=3)
  m.manual_seed(3)
  c = m.controlled_sample(emb, head, num_steps=9, eval_sp_size=5, sample_M=3)
  assert torch.equal(a, b) and torch.equal(a, c)
  cfg.time_conditioning = False


def test_build_eval_batches_and_baseline_rows_differ_across_shards(cuda):
  """F2: BaseModel.build_eval_batches (Enformer.py:135-160): per-timestep intermediate states
  and their final rewards.  And ADVICE r1: ddpm_cache rollouts honour row_offset, so two shards
  do not draw the same rows."""
  from svdd_b200.base_model import BaseModel
  torch.manual_seed(44)
  m = BaseModel(None, None, cdq=False, batch_size=6, val_batch_num=2, task='rna', random_init=True).to(cuda).eval()
  m.ref_model.config.sampling.steps = 8
  m.build_eval_batches()
  assert len(m.eval_time_step_batches) == 8 and len(m.eval_time_step_targets) == 8
  assert all(b.shape == (12, 50) for b in m.eval_time_step_batches)
  assert all(t.shape == (12,) and torch.equal(t, m.eval_time_step_targets[0]) for t in m.eval_time_step_targets)
  masked = [int((b == 4).sum()) for b in m.eval_time_step_batches]
  assert masked == sorted(masked, reverse=True) and masked[-1] == 0       # states get less masked; last one is clean
  m.ref_model.config.sampling.steps = 128
  d = m.ref_model
  d.sampler = 'ddpm_cache'
  d.manual_seed(9)
  a = d.decode_sample(num_steps=16, eval_sp_size=4, row_offset=0)
  d.manual_seed(9)
  b = d.decode_sample(num_steps=16, eval_sp_size=4, row_offset=4)
  d.manual_seed(9)
  w = d.decode_sample(num_steps=16, eval_sp_size=8, row_offset=0)
  assert not torch.equal(a, b)
  d.sampler = 'ddpm'


def test_cli_loads_reference_layout_value_checkpoint(cuda, tmp_path):
  """decode.py --load_checkpoint_path with a value-function .pt in the REFERENCE's key layout
  (reward_model.model.*, Lightning metric buffers; decode.py:101-104 strict=True)."""
  import subprocess, sys
  from svdd_b200.base_model import BaseModel
  torch.manual_seed(3)
  src = BaseModel(None, None, cdq=False, batch_size=8, val_batch_num=1, task='rna', random_init=True)
  sd = {}
  for k, v in src.state_dict().items():
    sd[('reward_model.model.' + k[len('reward_model.'):]) if k.startswith('reward_model.') else k] = v
  sd['reward_model.val_metrics.mse.total'] = torch.zeros(1)
  path = tmp_path / 'value.pt'
  torch.save({'epoch': 1, 'model_state_dict': sd}, path)
  cmd = [sys.executable, 'decode.py', '--task', 'rna', '--sample_M', '3', '--batch_size', '8', '--val_batch_num', '1',
         '--reward_name', 'MRL', '--random_init', '--out_dir', str(tmp_path), '--load_checkpoint_path', str(path)]
  r = subprocess.run(cmd, cwd=helpers.ROOT, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stdout + r.stderr
  assert np.load(tmp_path / 'rna-MRL.npz')['decoding'].shape == (8,)


def test_shard_equivalence_dna_headline_networks(cuda):
  """The same property on the headline configuration's networks (fused L = 200 denoiser, full
  Enformer value net, M = 10): a batch of 8 decoded whole == two shards of 4 with their row
  offsets, eager and through the graph.  Requires per-row results of every kernel to be
  independent of the batch composition (tile placement, chunking)."""
  m = _dna_model(cuda)
  emb, head = helpers.build_enformer(full=True)
  emb, head = emb.to(cuda), head.to(cuda)
  for graph in (False, True):
    m.use_cuda_graph = graph

    def run(rows, off):
      m.manual_seed(31)
      return m.controlled_sample(emb, head, num_steps=3, eval_sp_size=rows, sample_M=10, row_offset=off)
    whole = run(8, 0)
    assert torch.equal(whole, torch.cat([run(4, 0), run(4, 4)], 0)), f'graph={graph}'
    assert torch.equal(whole, torch.cat([run(5, 0), run(3, 5)], 0)), f'graph={graph}'


@pytest.mark.parametrize('script,extra,name', [('decode.py', [], 'dna-HepG2.npz'),
                                               ('decode_tweedie.py', ['--tweedie', 'True'], 'dna-HepG2_tw.npz')])
def test_cli_dna_decode_and_tweedie(cuda, tmp_path, script, extra, name):
  """BASELINE configs 2 / 3 through the reference's CLI surface: decode.py --task dna (SVDD-MC,
  Enformer value net built as decode.py:78-80 does) and decode_tweedie.py --tweedie True (SVDD-PM,
  3-task Enformer reward oracle), random-init weights, small batch: ./log/dna-HepG2[_tw].npz with
  float32 (N,) arrays `decoding` and `baseline`."""
  import subprocess, sys
  cmd = [sys.executable, script, '--task', 'dna', '--sample_M', '3', '--batch_size', '8', '--val_batch_num', '1',
         '--reward_name', 'HepG2', '--random_init', '--out_dir', str(tmp_path)] + extra
  r = subprocess.run(cmd, cwd=helpers.ROOT, capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  d = np.load(tmp_path / name)
  assert sorted(d.files) == ['baseline', 'decoding']
  assert d['decoding'].shape == (8,) and d['decoding'].dtype == np.float32
  assert d['baseline'].shape == (8,) and np.isfinite(d['decoding']).all() and np.isfinite(d['baseline']).all()


@pytest.mark.parametrize('task,tweedie,alpha', [('rna', False, 0.0), ('rna', True, 0.0), ('rna', False, 0.1),
                                                ('dna', False, 0.0), ('dna', True, 0.0)])
def test_svdd_step_one_call_equals_the_stage_calls(cuda, task, tweedie, alpha):
  """svdd_step (include/svdd_b200.h: the whole body of _ddpm_update_finetune_controlled /
  _controlled_twedie behind ONE C call) is bit-identical to sequencing the stage entry points,
  for both scorer kinds, MC and PM, argmax and soft selection, with Philox and injected noise."""
  m = _rna_model(cuda) if task == 'rna' else _dna_model(cuda)
  L = 50 if task == 'rna' else 200
  B, M = (6, 5) if task == 'rna' else (4, 10)
  if task == 'rna':
    emb, head = helpers.build_convgru_oracle() if tweedie else helpers.build_convgru_value()
  else:
    emb, head = helpers.build_enformer(full=False)
  scorer = diffusion_gosai._as_scorer(emb.to(cuda), head.to(cuda))
  den = m.backbone.packed()
  g = torch.Generator().manual_seed(77)
  x = torch.randint(0, 4, (B, L), generator=g).to(torch.uint8)
  x[torch.rand(B, L, generator=g) < 0.6] = 4
  x = x.to(cuda)
  mc_t, mc_s = 0.61, 0.55
  for U in (None, torch.rand(M, B, L, 5, generator=g).to(cuda)):
    kw = dict(U=U, seed=991, step=7, row_offset=3)
    logits = den.forward(x, 0.0)
    cand, q = _lib.subs_sample(logits, x, M, mc_t, mc_s, want_q=True, **kw)
    flat = cand.reshape(M * B, L)
    if tweedie:
      flat = _lib.x0_argmax(den.forward(flat, 0.0), flat)
    scores = scorer.score(flat).reshape(M, B)
    x_ref, idx_ref = _lib.select_gather(scores, cand, alpha=alpha, seed=991, step=7, row_offset=3, want_idx=True)
    before = _lib.launch_count()
    got = _lib.step(den, scorer, x, M, mc_t, mc_s, tweedie=tweedie, alpha=alpha, want_q=True, **kw)
    assert _lib.launch_count() > before
    assert torch.equal(got['cand'], cand)
    assert torch.equal(got['q'], q)
    assert torch.equal(got['scores'], scores)
    assert torch.equal(got['idx'], idx_ref)
    assert torch.equal(got['x'], x_ref)
    # carry-over: an unmasked position never changes (copy_flag of the reference's return tuple)
    keep = x != 4
    assert torch.equal(got['x'][keep], x[keep])


def test_svdd_step_is_graph_capturable_and_rejects_a_short_workspace(cuda):
  m = _rna_model(cuda)
  emb, head = helpers.build_convgru_value()
  scorer = diffusion_gosai._as_scorer(emb.to(cuda), head.to(cuda))
  den = m.backbone.packed()
  B, M = 8, 4
  x = torch.full((B, 50), 4, dtype=torch.uint8, device=cuda)
  seed_dev = torch.tensor([5, 0], dtype=torch.int64, device=cuda)
  eager = _lib.step(den, scorer, x, M, 0.9, 0.8, seed_dev=seed_dev)
  out = torch.empty_like(x)
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.stream(s):
    with torch.cuda.graph(graph, stream=s):
      held = _lib.step(den, scorer, x, M, 0.9, 0.8, seed_dev=seed_dev, out=out, ws=eager['ws'])
  graph.replay()
  torch.cuda.synchronize()
  assert torch.equal(out, eager['x'])
  seed_dev[0] = 6                      # a new run key replays the same graph with fresh noise
  graph.replay()
  torch.cuda.synchronize()
  assert not torch.equal(out, eager['x'])
  del held
  # too small a workspace is an error, not a write past the end
  import ctypes
  a = _lib._StepArgs()
  a.denoiser, a.scorer, a.scorer_kind = den._h, scorer._h, 0
  tb = den.time_bias(0.0)
  a.time_bias_t, a.x, a.x_out, a.tok_dtype = tb.data_ptr(), x.data_ptr(), out.data_ptr(), _lib.SVDD_TOK_U8
  a.B, a.L, a.M, a.mc_t, a.mc_s = B, 50, M, 0.9, 0.8
  need = _lib.lib().svdd_step_workspace_bytes(ctypes.byref(a))
  assert need > 0
  ws = torch.empty(need, dtype=torch.uint8, device=cuda)
  a.ws, a.ws_bytes = ws.data_ptr(), need - 256
  assert _lib.lib().svdd_step(ctypes.byref(a)) == -4
  assert b'workspace' in _lib.lib().svdd_last_error()
