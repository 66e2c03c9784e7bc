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
