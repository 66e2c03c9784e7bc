"""CPU restatement of the SVDD sampler (stages 2 and 4 and the outer loops) --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch CPU fp32, written to follow the reference's operator order so that, with
the same weights and the same uniform-noise tensors, token outputs equal the
reference's bit-for-bit on CPU (tests/golden/*.npz pin this).

Noise is always explicit: every function that draws takes the uniform tensor
``U`` it should use.  ``TorchNoise`` reproduces the reference's RNG consumption
order (one ``torch.rand_like(q_xs)`` per candidate, m-major:
diffusion_gosai.py:1203) so that seeding torch identically reproduces a
reference run; ``PhiloxNoise`` reproduces the in-kernel counter-based stream of
``svdd_subs_sample`` when it is called with ``U == NULL``.
"""
import numpy as np
import torch

from . import philox

MASK_INDEX = 4            # diffusion_gosai.py:85,94-95 (vocab 4 + mask)
NEG_INFINITY = -1000000.0  # diffusion_gosai.py:156


# ----------------------------------------------------------------------------
# A3: noise schedule (noise_schedule.py:126-145; use at diffusion_gosai.py:1176-1187)
# ----------------------------------------------------------------------------

def total_noise(t, eps=1e-3):
  """LogLinearNoise.total_noise: sigma(t) = -log1p(-(1-eps) t)."""
  return -torch.log1p(-(1 - eps) * t)


def move_chances(num_steps=128, eps=1e-5):
  """Per-step (mc_t, mc_s, sigma_t, sigma_s) as fp32 scalars, computed with the
  reference's exact expression sequence (diffusion_gosai.py:1036-1043 and
  1176-1187) on a [1,1] tensor: t = timesteps[i]*ones; s = t - dt;
  mc = 1 - exp(-sigma).  Also returns sigma at timesteps[-1] for the
  noise-removal forward (:1049-1054)."""
  timesteps = torch.linspace(1, eps, num_steps + 1)
  dt = (1 - eps) / num_steps
  rows = []
  for i in range(num_steps):
    t = timesteps[i] * torch.ones(1, 1)
    sigma_t = total_noise(t).squeeze(-1)
    sigma_s = total_noise(t - dt).squeeze(-1)
    mc_t = 1 - torch.exp(-sigma_t)
    mc_s = 1 - torch.exp(-sigma_s)
    rows.append((mc_t.item(), mc_s.item(), sigma_t.item(), sigma_s.item()))
  sched = np.asarray(rows, dtype=np.float32)
  sigma_last = total_noise(timesteps[-1] * torch.ones(1, 1)).item()
  return sched, np.float32(sigma_last)


# ----------------------------------------------------------------------------
# A6-A8: SUBS parameterisation, q_xs, Gumbel-max draw, carry-over
# ----------------------------------------------------------------------------

def subs_parameterization(logits, xt):
  """Diffusion._subs_parameterization (diffusion_gosai.py:286-304).
  logits fp32[B,L,5] (not modified), xt int64[B,L] -> log_p fp32[B,L,5]."""
  logits = logits.clone()
  logits[:, :, MASK_INDEX] += NEG_INFINITY
  logits = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
  unmasked = xt != MASK_INDEX
  logits[unmasked] = NEG_INFINITY
  logits[unmasked, xt[unmasked]] = 0
  return logits


def build_q_xs(log_p, mc_t, mc_s):
  """diffusion_gosai.py:1194-1196: q = exp(log_p)*(mc_t-mc_s); q[...,4]=mc_s.
  mc_t / mc_s are fp32 python/numpy scalars (identical across the batch)."""
  mc_t = torch.tensor(mc_t, dtype=torch.float32)
  mc_s = torch.tensor(mc_s, dtype=torch.float32)
  q = log_p.exp() * (mc_t - mc_s)
  q[:, :, MASK_INDEX] = mc_s
  return q


def gumbel_argmax(q, U):
  """_sample_categorical (diffusion_gosai.py:30-34) with the uniform tensor
  made explicit: argmax(q / (1e-10 - log(U + 1e-10)), -1), first index on
  ties."""
  g = 1e-10 - (U + 1e-10).log()
  return (q / g).argmax(dim=-1)


def draw_candidates(x, q, U):
  """diffusion_gosai.py:1199-1203.  U fp32[M,B,L,5] -> candidates int64[M,B,L];
  already-unmasked tokens are carried over regardless of the draw."""
  copy_flag = (x != MASK_INDEX).to(x.dtype)
  return torch.stack([copy_flag * x + (1 - copy_flag) * gumbel_argmax(q, U[m])
                      for m in range(U.shape[0])], dim=0)


def draw_margin(q, U):
  """Relative gap between the best and second-best Gumbel-max key of every
  draw; a test may exempt draws whose margin is within a few fp32 ulps when it
  compares against a different libm (CPU SLEEF vs CUDA libdevice)."""
  g = 1e-10 - (U + 1e-10).log()
  top2 = (q / g).topk(2, dim=-1).values
  return (top2[..., 0] - top2[..., 1]) / top2[..., 0].clamp_min(1e-30)


def transform_samples(samples, num_classes=4):
  """Diffusion.transform_samples (diffusion_gosai.py:1462-1470): one-hot(4)
  with mask rows all-zero.  int64[B,L] -> int64[B,L,4]."""
  keep = samples != MASK_INDEX
  oh = torch.nn.functional.one_hot(samples * keep, num_classes=num_classes)
  return oh * keep.unsqueeze(-1)


# ----------------------------------------------------------------------------
# A14: selection
# ----------------------------------------------------------------------------

def select(scores, alpha=0.0, U=None):
  """diffusion_gosai.py:1219-1225 for alpha == 0: softmax over the M scores of
  each sequence, THEN argmax (first index on ties; the softmax matters because
  exp(s - max) collapses near-equal scores).

  alpha > 0 is NOT reference behaviour (the reference's multinomial lines are
  commented out, :1213,1223).  Build decision, following the SVDD paper's soft
  rule and reusing the reference's own draw primitive:
      idx = _sample_categorical(softmax(scores / alpha, dim=1))   with U[B,M].
  scores fp32[B,M] -> idx int64[B].
  """
  if alpha == 0.0:
    return torch.softmax(scores, dim=1).argmax(dim=1)
  p = torch.softmax(scores / torch.tensor(alpha, dtype=torch.float32), dim=1)
  return gumbel_argmax(p, U)


def gather_selected(cand, idx):
  """diffusion_gosai.py:1226-1227: x_next[b] = cand[idx[b], b, :]."""
  B = cand.shape[1]
  return cand[idx, torch.arange(B)]


# ----------------------------------------------------------------------------
# Noise providers
# ----------------------------------------------------------------------------

class TorchNoise:
  """Consumes the global torch CPU generator exactly like the reference.

  The reference's q_xs inherits the memory layout of CNNModel's output, which
  is a [B,5,L] buffer viewed as [B,L,5] (``feat.permute(0, 2, 1)``,
  models/dnaconv.py:202); ``torch.rand_like`` preserves those strides and fills
  in MEMORY order, so the stream lands as rand(B,5,L).permute(0,2,1)."""

  def draws(self, step, M, B, L, V=5):
    return torch.stack([torch.rand(B, V, L).permute(0, 2, 1) for _ in range(M)],
                       dim=0)

  def select(self, step, B, M):
    return torch.rand(B, M)


class ArrayNoise:
  """Injected noise: U_draw[step] -> [M,B,L,5]; U_sel[step] -> [B,M]."""

  def __init__(self, U_draw, U_sel=None):
    self.U_draw, self.U_sel = U_draw, U_sel

  def draws(self, step, M, B, L, V=5):
    return torch.as_tensor(self.U_draw[step])

  def select(self, step, B, M):
    return torch.as_tensor(self.U_sel[step])


class PhiloxNoise:
  """The counter-based stream of the CUDA kernels (see oracle/philox.py)."""

  def __init__(self, seed, row_offset=0):
    self.seed, self.row_offset = seed, row_offset

  def draws(self, step, M, B, L, V=5):
    return torch.from_numpy(philox.draw_uniforms(
        self.seed, step, M, B, L, self.row_offset))

  def select(self, step, B, M):
    return torch.from_numpy(philox.select_uniforms(
        self.seed, step, B, M, self.row_offset))


# ----------------------------------------------------------------------------
# A1, A16-A18: steps and outer loops
# ----------------------------------------------------------------------------

def forward_log_p(denoiser, x):
  """Diffusion.forward (diffusion_gosai.py:339-357) with time_conditioning
  False: backbone(x, 0) then SUBS."""
  return subs_parameterization(denoiser(x), x)


def step_plain(denoiser, x, mc_t, mc_s, U):
  """_ddpm_update_finetune (diffusion_gosai.py:1148-1172), M = 1."""
  q = build_q_xs(forward_log_p(denoiser, x), mc_t, mc_s)
  return draw_candidates(x, q, U[None] if U.dim() == 3 else U)[0]


def step_mc(denoiser, value_fn, x, mc_t, mc_s, U, alpha=0.0, U_sel=None,
            trace=None):
  """_ddpm_update_finetune_controlled (diffusion_gosai.py:1175-1228).
  value_fn: int64[B,L] candidate tokens -> fp32[B] (it applies
  transform_samples + embedding + head + squeeze, :1208-1209)."""
  logits = denoiser(x)
  q = build_q_xs(subs_parameterization(logits, x), mc_t, mc_s)
  cand = draw_candidates(x, q, U)
  scores = torch.stack([value_fn(cand[m]) for m in range(cand.shape[0])], dim=1)
  idx = select(scores, alpha, U_sel)
  x_next = gather_selected(cand, idx)
  if trace is not None:
    trace.append(dict(x=x, logits=logits, q=q, cand=cand, scores=scores,
                      idx=idx, x_next=x_next))
  return x_next


def tweedie_onehot(denoiser, cand_m):
  """diffusion_gosai.py:1415-1419 (options == "True"): argmax of the post-SUBS
  log-probs of the candidate, one-hot, carried tokens kept.  Returns the x0
  token estimate int64[B,L] (values 0..3) -- its one-hot(4) is what the
  reward oracle sees."""
  expected = forward_log_p(denoiser, cand_m)
  return expected.argmax(dim=2)


def step_pm(denoiser, reward_fn, x, mc_t, mc_s, U, tweedie=True, alpha=0.0,
            U_sel=None, trace=None):
  """_ddpm_update_finetune_controlled_twedie (diffusion_gosai.py:1374-1460).
  reward_fn: int64[B,L] tokens in 0..4 (4 -> all-zero one-hot row) -> fp32[B]
  (reward_model(onehot.float().transpose(1,2))[:,0].squeeze(), :1430-1436)."""
  logits = denoiser(x)
  q = build_q_xs(subs_parameterization(logits, x), mc_t, mc_s)
  cand = draw_candidates(x, q, U)
  scores = []
  for m in range(cand.shape[0]):
    toks = tweedie_onehot(denoiser, cand[m]) if tweedie else cand[m]
    scores.append(reward_fn(toks))
  scores = torch.stack(scores, dim=1)
  idx = select(scores, alpha, U_sel)
  x_next = gather_selected(cand, idx)
  if trace is not None:
    trace.append(dict(x=x, logits=logits, q=q, cand=cand, scores=scores,
                      idx=idx, x_next=x_next))
  return x_next


def noise_removal(denoiser, x):
  """diffusion_gosai.py:1049-1060: x = forward(x)[:, :, :-1].argmax(-1)."""
  return forward_log_p(denoiser, x)[:, :, :-1].argmax(dim=-1)


def _loop(step_fn, B, L, num_steps, eps, noise, M, alpha, denoiser,
          noise_removal_on=True):
  sched, _ = move_chances(num_steps, eps)
  x = torch.full((B, L), MASK_INDEX, dtype=torch.int64)   # _sample_prior :751
  for i in range(num_steps):
    U = noise.draws(i, M, B, L)
    U_sel = noise.select(i, B, M) if alpha > 0 else None
    x = step_fn(x, float(sched[i, 0]), float(sched[i, 1]), U, U_sel)
  if noise_removal_on:
    x = noise_removal(denoiser, x)
  return x


def controlled_sample(denoiser, value_fn, B, L, M=10, num_steps=128, eps=1e-5,
                      noise=None, alpha=0.0, trace=None):
  """Diffusion.controlled_sample (diffusion_gosai.py:1022-1061) -- SVDD-MC."""
  noise = noise or TorchNoise()
  return _loop(lambda x, a, b, U, Us: step_mc(denoiser, value_fn, x, a, b, U,
                                              alpha, Us, trace),
               B, L, num_steps, eps, noise, M, alpha, denoiser)


def controlled_sample_tweedie(denoiser, reward_fn, B, L, M=10, num_steps=128,
                              eps=1e-5, noise=None, tweedie=True, alpha=0.0,
                              trace=None):
  """Diffusion.controlled_sample_tweedie (diffusion_gosai.py:1106-1145) --
  SVDD-PM."""
  noise = noise or TorchNoise()
  return _loop(lambda x, a, b, U, Us: step_pm(denoiser, reward_fn, x, a, b, U,
                                              tweedie, alpha, Us, trace),
               B, L, num_steps, eps, noise, M, alpha, denoiser)


def decode_sample(denoiser, B, L, num_steps=128, eps=1e-5, noise=None):
  """Diffusion.decode_sample (diffusion_gosai.py:889-936), predictor 'ddpm'."""
  noise = noise or TorchNoise()
  return _loop(lambda x, a, b, U, Us: step_plain(denoiser, x, a, b, U),
               B, L, num_steps, eps, noise, 1, 0.0, denoiser)


def sample_with_mid(denoiser, B, L, num_steps=128, eps=1e-5, noise=None):
  """Diffusion._sample (diffusion_gosai.py:821-886), cdq False: also returns
  the num_steps-1 intermediate states (feeds value-function training)."""
  noise = noise or TorchNoise()
  sched, _ = move_chances(num_steps, eps)
  x = torch.full((B, L), MASK_INDEX, dtype=torch.int64)
  mid = []
  for i in range(num_steps):
    x = step_plain(denoiser, x, float(sched[i, 0]), float(sched[i, 1]),
                   noise.draws(i, 1, B, L))
    if i != num_steps - 1:
      mid.append(x.clone())
  return noise_removal(denoiser, x), mid


def sample_ddpm_cache(denoiser, B, L, num_steps=128, eps=1e-5, noise=None, time_conditioning=False,
                      log_p_fn=None):
  """Diffusion._sample with sampling.predictor == 'ddpm_cache' (diffusion_gosai.py:755-773,
  :858-865): the move chances are t and t - dt themselves (no noise schedule), and the
  post-SUBS probabilities p_x0 are reused for the next step while the whole batch is
  unchanged (and time conditioning is off).  Returns (x, mid, number of denoiser forwards
  incl. noise removal).  `log_p_fn(x)` may replace SUBS(denoiser(x)) (the GPU tests feed
  the engine's own log-probabilities)."""
  noise = noise or TorchNoise()
  fwd = log_p_fn if log_p_fn is not None else (lambda x: forward_log_p(denoiser, x))
  ts = torch.linspace(1, eps, num_steps + 1)              # :835-836
  dt = (1 - eps) / num_steps                                # :837
  x = torch.full((B, L), MASK_INDEX, dtype=torch.int64)
  mid, p_cache, n_fwd = [], None, 0
  for i in range(num_steps):
    t = ts[i] * torch.ones(B, 1)
    mc_t = t[:, None, :]                                    # :761 move_chance_t = t[:, None, None]
    mc_s = (t - dt)[:, None, :]                             # :762
    if p_cache is None:
      p_cache = fwd(x).exp()                                # :765
      n_fwd += 1
    q = p_cache * (mc_t - mc_s)                             # :768
    q[:, :, MASK_INDEX] = mc_s[:, :, 0]                     # :769
    x_next = draw_candidates(x, q, noise.draws(i, 1, B, L))[0]
    if not torch.equal(x_next, x) or time_conditioning:     # :861-864 (allclose on integers)
      p_cache = None
    x = x_next
    if i != num_steps - 1:
      mid.append(x.clone())
  x = fwd(x)[:, :, :-1].argmax(dim=-1)                      # noise removal :872-880
  return x, mid, n_fwd + 1
