"""-m gpu: the CUDA stage-2 / stage-4 kernels (through the C ABI) against the
reference goldens and the CPU oracle.  Integer outputs must be bit-exact.

CPU (SLEEF) and CUDA (libdevice) logf/expf may differ in the last ulp, so a draw
whose two best Gumbel keys are within a few ulps of each other can legitimately
flip; such draws are identified with oracle.svdd.draw_margin and must be (a)
absent at fixture sizes and (b) the ONLY mismatches at large sizes."""
import numpy as np
import pytest
import torch

import helpers
from oracle import philox, svdd
from svdd_b200 import _lib

pytestmark = pytest.mark.gpu
NEAR_TIE = 8 * 2.0 ** -23


def T(a):
  return torch.from_numpy(np.asarray(a))


def _cases():
  g = helpers.load_golden('stage_kats.npz')
  for tag in 'abcd':
    yield tag, {k[2:]: g[k] for k in g.files if k.startswith(tag + '_')}


@pytest.mark.parametrize('tok_dtype', [torch.int64, torch.uint8])
def test_stage2_golden(cuda, tok_dtype):
  sched, _ = svdd.move_chances(128, 1e-5)
  for tag, c in _cases():
    step = int(c['step'])
    x = T(c['x']).to(cuda).to(tok_dtype)
    logits, U = T(c['logits']).to(cuda), T(c['U']).to(cuda)
    M = U.shape[0]
    cand, q = _lib.subs_sample(logits, x, M, sched[step, 0], sched[step, 1], U=U, want_q=True)
    assert cand.dtype == tok_dtype
    np.testing.assert_array_equal(cand.cpu().numpy().astype(np.int64), c['cand'], err_msg=tag)
    # q_xs: same IEEE ops; CPU-vs-CUDA libm may differ by 1 ulp in logsumexp (values
    # of a few units -> ~5e-7 absolute in log_p -> same relative in exp(log_p))
    np.testing.assert_allclose(q.cpu().numpy(), c['q'], rtol=2e-6, atol=1e-37, err_msg=tag)
    # the post-SUBS entry point (is_log_p) gives the same candidates
    cand2 = _lib.subs_sample(T(c['log_p']).to(cuda), x, M, sched[step, 0], sched[step, 1], U=U,
                             is_log_p=True)
    assert torch.equal(cand2, cand)
    x0 = _lib.x0_argmax(logits, x)
    np.testing.assert_array_equal(x0.cpu().numpy().astype(np.int64), c['x0'], err_msg=tag)


@pytest.mark.parametrize('tok_dtype', [torch.int64, torch.uint8])
def test_stage4_golden(cuda, tok_dtype):
  for tag, c in _cases():
    cand = T(c['cand']).to(cuda).to(tok_dtype)
    scores = T(c['scores']).t().contiguous().to(cuda)        # [B,M] -> [M,B]
    x_next, idx = _lib.select_gather(scores, cand, want_idx=True)
    ref_idx = svdd.select(T(c['scores']))
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx.numpy(), err_msg=tag)
    np.testing.assert_array_equal(x_next.cpu().numpy().astype(np.int64), c['x_next'], err_msg=tag)


def test_stage2_edge_cases(cuda):
  sched, _ = svdd.move_chances(128, 1e-5)
  # empty batch, ragged sizes that do not fill a warp / block, M = 1
  for B, L, M in [(0, 50, 3), (1, 1, 1), (3, 33, 2), (7, 200, 1), (2, 257, 5)]:
    g = torch.Generator().manual_seed(B * 1000 + L)
    logits = torch.randn(B, L, 5, generator=g)
    x = helpers.random_tokens(B, L, 5, 0.5) if B else torch.zeros((0, L), dtype=torch.int64)
    U = torch.rand(M, B, L, 5, generator=g)
    cand = _lib.subs_sample(logits.to(cuda), x.to(cuda), M, sched[3, 0], sched[3, 1], U=U.to(cuda))
    q = svdd.build_q_xs(svdd.subs_parameterization(logits, x), sched[3, 0], sched[3, 1])
    ref = svdd.draw_candidates(x, q, U) if B else torch.zeros((M, 0, L), dtype=torch.int64)
    assert torch.equal(cand.cpu(), ref), (B, L, M)
  # all-mask prior at step 0; last step (mc_s ~ 1e-5); U == 0 -> argmax(q)
  x = torch.full((4, 50), 4, dtype=torch.int64)
  logits = torch.randn(4, 50, 5, generator=torch.Generator().manual_seed(1))
  for step in (0, 127):
    q = svdd.build_q_xs(svdd.subs_parameterization(logits, x), sched[step, 0], sched[step, 1])
    cand = _lib.subs_sample(logits.to(cuda), x.to(cuda), 2, sched[step, 0], sched[step, 1],
                            U=torch.zeros(2, 4, 50, 5, device=cuda))
    assert torch.equal(cand[0].cpu(), q.argmax(-1)) and torch.equal(cand[1], cand[0])


def test_stage2_large_property(cuda):
  """BASELINE config-4 sized step (B=4096/8 per GPU, L=200, M=20): carried tokens
  are untouched, draws agree with the oracle except at provable near-ties."""
  sched, _ = svdd.move_chances(128, 1e-5)
  B, L, M, step = 512, 200, 20, 64
  g = torch.Generator().manual_seed(9)
  logits = torch.randn(B, L, 5, generator=g) * 2
  x = helpers.random_tokens(B, L, 10, 0.5)
  U = torch.rand(M, B, L, 5, generator=g)
  cand = _lib.subs_sample(logits.to(cuda), x.to(cuda), M, sched[step, 0], sched[step, 1],
                          U=U.to(cuda)).cpu()
  keep = x != 4
  assert torch.equal(cand[:, keep], x[keep].expand(M, -1))
  assert int(cand[:, ~keep].max()) <= 4
  q = svdd.build_q_xs(svdd.subs_parameterization(logits, x), sched[step, 0], sched[step, 1])
  ref = svdd.draw_candidates(x, q, U)
  bad = cand != ref
  if bad.any():
    margin = torch.stack([svdd.draw_margin(q, U[m]) for m in range(M)])
    assert float(margin[bad].max()) < NEAR_TIE, 'mismatch that is not a near-tie'
  assert int(bad.sum()) <= 4, int(bad.sum())


@pytest.mark.parametrize('tok_dtype', [torch.int64, torch.uint8])
def test_stage2_fast_path_equals_exact_path(cuda, tok_dtype):
  """svdd_subs_sample decides a draw with approximate logs only when an error-bound test
  proves the exact argmax; it must therefore agree bit for bit with the all-exact kernel --
  on random inputs at BASELINE config-4 per-GPU size and on adversarial inputs built to sit
  on ties and near-ties (equal keys, keys 1 ulp apart, u next to 0 and to 1)."""
  sched, _ = svdd.move_chances(128, 1e-5)
  g = torch.Generator().manual_seed(123)
  B, L, M = 512, 200, 20
  logits = (torch.randn(B, L, 5, generator=g) * 3).to(cuda)
  x = helpers.random_tokens(B, L, 10, 0.7).to(cuda).to(tok_dtype)
  U = torch.rand(M, B, L, 5, generator=g).to(cuda)
  for step in (0, 64, 127):
    a = _lib.subs_sample(logits, x, M, sched[step, 0], sched[step, 1], U=U)
    b = _lib.subs_sample(logits, x, M, sched[step, 0], sched[step, 1], U=U, exact=True)
    assert torch.equal(a, b), step
  # adversarial: flat logits (equal q over the 4 bases) and noise with repeated / adjacent values
  B, L, M = 64, 200, 16
  x = torch.full((B, L), 4, dtype=tok_dtype, device=cuda)
  logits = torch.zeros(B, L, 5, device=cuda)
  base = torch.rand(M, B, L, 1, generator=g).expand(M, B, L, 5).clone()
  ulp = torch.tensor(2.0 ** -24)
  U = base.clone()
  U[..., 1] = torch.nextafter(base[..., 1], torch.tensor(1.0))        # 1 ulp above
  U[..., 2] = torch.nextafter(base[..., 2], torch.tensor(0.0))        # 1 ulp below
  U[0::4, ..., :] = base[0::4]                                        # exact ties -> first index
  U[1::4, ..., 3] = 1.0 - ulp                                         # g ~ 6e-8: tiny denominators
  U[2::4, ..., 0] = 0.0                                               # g = 23.03: largest denominator
  U[3::4, ..., 4] = torch.rand(M // 4, B, L, generator=g) * 1e-6 + (1.0 - 1e-6 - ulp)
  U = U.clamp_(0.0, float(1.0 - ulp)).to(cuda)
  for step in (0, 100, 127):
    a = _lib.subs_sample(logits, x, M, sched[step, 0], sched[step, 1], U=U)
    b = _lib.subs_sample(logits, x, M, sched[step, 0], sched[step, 1], U=U, exact=True)
    assert torch.equal(a, b), step
  # exact ties resolve to the first index, as torch.argmax does
  assert int(a[0].max()) == 0


def test_stage2_philox_stream(cuda):
  """U == NULL: the in-kernel Philox stream equals oracle/philox.py, is
  independent of how rows are sharded, and differs per step / seed."""
  sched, _ = svdd.move_chances(128, 1e-5)
  B, L, M, step, seed = 6, 50, 4, 17, 0x1234567890ABCDEF
  logits = torch.randn(B, L, 5, generator=torch.Generator().manual_seed(3))
  x = helpers.random_tokens(B, L, 4, 0.8)
  U = torch.from_numpy(philox.draw_uniforms(seed, step, M, B, L))
  q = svdd.build_q_xs(svdd.subs_parameterization(logits, x), sched[step, 0], sched[step, 1])
  ref = svdd.draw_candidates(x, q, U)
  cand = _lib.subs_sample(logits.to(cuda), x.to(cuda), M, sched[step, 0], sched[step, 1],
                          seed=seed, step=step)
  assert torch.equal(cand.cpu(), ref)
  lo = _lib.subs_sample(logits[:2].to(cuda), x[:2].to(cuda), M, sched[step, 0], sched[step, 1],
                        seed=seed, step=step, row_offset=0)
  hi = _lib.subs_sample(logits[2:].to(cuda), x[2:].to(cuda), M, sched[step, 0], sched[step, 1],
                        seed=seed, step=step, row_offset=2)
  assert torch.equal(torch.cat([lo, hi], 1), cand)
  other = _lib.subs_sample(logits.to(cuda), x.to(cuda), M, sched[step, 0], sched[step, 1],
                           seed=seed, step=step + 1)
  assert not torch.equal(other, cand)


@pytest.mark.parametrize('M', [1, 2, 10, 20, 50, 64])
def test_stage4_properties(cuda, M):
  B, L = 37, 50
  g = torch.Generator().manual_seed(M)
  cand = torch.randint(0, 5, (M, B, L), generator=g)
  for scale in (1.0, 1e-6, 30.0):
    scores = torch.randn(B, M, generator=g) * scale
    x_next, idx = _lib.select_gather(scores.t().contiguous().to(cuda), cand.to(cuda), want_idx=True)
    ref = svdd.select(scores)
    sm = torch.softmax(scores, 1)
    bad = idx.cpu().long() != ref
    # a mismatch is only tolerated between candidates whose softmax values are within 2 ulp
    if bad.any():
      got = sm[bad, idx.cpu().long()[bad]]
      want = sm[bad, ref[bad]]
      assert float(((want - got).abs() / want).max()) < 3e-7
    assert torch.equal(x_next.cpu(), svdd.gather_selected(cand, idx.cpu().long()))
  # alpha > 0 with injected and Philox uniforms
  scores = torch.randn(B, M, generator=g)
  U = torch.rand(B, M, generator=g)
  for alpha in (0.1, 1.0):
    _, idx = _lib.select_gather(scores.t().contiguous().to(cuda), cand.to(cuda), alpha=alpha,
                                U_sel=U.to(cuda), want_idx=True)
    ref = svdd.select(scores, alpha, U)
    assert (idx.cpu().long() != ref).sum() <= 1
    _, idx = _lib.select_gather(scores.t().contiguous().to(cuda), cand.to(cuda), alpha=alpha,
                                seed=77, step=5, row_offset=3, want_idx=True)
    Up = torch.from_numpy(philox.select_uniforms(77, 5, B, M, row_offset=3))
    assert (idx.cpu().long() != svdd.select(scores, alpha, Up)).sum() <= 1


def test_no_cpu_fallback():
  with pytest.raises(_lib.SvddError):
    _lib.subs_sample(torch.zeros(1, 2, 5), torch.zeros(1, 2, dtype=torch.int64), 1, 0.5, 0.4)


def test_schedule_host_vs_device(cuda):
  """the product evaluates the move-chance schedule with CPU libm (what the oracle and
  the reference-generated goldens pin); the reference's decode path, run on a GPU, evaluates the
  same expression sequence with CUDA libm (diffusion_gosai.py:1036-1043, 1176-1187).  Measured
  for every schedule length the tests and BASELINE configs use: about a quarter of the values
  differ, always by ONE ulp of exp(-sigma) (<= 2^-23 absolute on the move chance; after the
  cancellation in 1 - exp(-sigma) that is up to 2^16 ulps of the LAST step's mc_s ~ 1e-5).  A
  difference can only flip a Gumbel draw whose two best keys are within that ulp
  (oracle.svdd.draw_margin identifies such draws; none occur in the fixtures).
  ``Diffusion.schedule_device = 'model'`` evaluates on the model's device instead, for bitwise
  agreement with a reference run on the same GPU."""
  from svdd_b200 import config, diffusion_gosai, noise_schedule
  noise = noise_schedule.LogLinearNoise()
  worst, differing, total = 0.0, 0, 0
  for steps in (3, 4, 6, 8, 12, 16, 128, 160):
    host, s_host = noise_schedule.move_chance_schedule(noise, steps, 1e-5)
    dev, s_dev = noise_schedule.move_chance_schedule(noise.to(cuda), steps, 1e-5, device=cuda)
    h = np.float32([r[:2] for r in host]).astype(np.float64)
    d = np.float32([r[:2] for r in dev]).astype(np.float64)
    worst = max(worst, float(np.abs(h - d).max()))
    differing += int((h != d).sum())
    total += h.size
  print(f'\n[schedule] host vs device move chances: {differing} of {total} values differ, worst |d| = {worst:.3e} '
        f'({worst * 2 ** 23:.2f} x 2^-23)')
  assert worst <= 2.0 ** -23
  # the switch: same trajectory machinery, schedule evaluated where the model lives
  torch.manual_seed(44)
  m = diffusion_gosai.Diffusion(config.load_config('rna')).to(cuda).eval()
  m.schedule_device = 'model'
  m.manual_seed(1)
  a = m.decode_sample(num_steps=8, eval_sp_size=4)
  m.schedule_device = 'cpu'
  m.manual_seed(1)
  b = m.decode_sample(num_steps=8, eval_sp_size=4)
  assert a.shape == b.shape == (4, 50) and int(a.max()) <= 3
  assert float((a != b).float().mean()) < 0.05        # 1-ulp move chances: (nearly always) the same tokens
