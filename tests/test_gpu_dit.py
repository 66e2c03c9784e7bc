"""-m gpu: the DiT denoiser backbone (models/dit.py, `backbone: dit`) through the C ABI
(svdd_dit_*) against the reference-module goldens and the oracle; then as stage 1 of the SVDD
loops.  Tolerances as for the CNN denoiser (tests/test_gpu_nets.py): max |d| <= TIGHT * scale vs the
oracle's bf16-operand emulation, <= LOOSE * scale vs the reference's fp32 output."""
import numpy as np
import pytest
import torch

import helpers
from oracle import nets, svdd
from svdd_b200 import _lib, diffusion_gosai, value_nets
from test_oracle_golden import DIT_CASES

pytestmark = pytest.mark.gpu

TIGHT, LOOSE = 6e-3, 3e-2


def T(a):
  return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('tag,nb,L', DIT_CASES)
def test_dit_logits(cuda, tag, nb, L):
  g = helpers.load_golden('dit_seed44.npz')
  m = helpers.build_dit(n_blocks=nb, length=L)
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  x, ref, sigma = T(g[f'{tag}_tokens']), T(g[f'{tag}_logits']), float(g[f'{tag}_sigma'])
  with torch.no_grad():
    emu = nets.dit_logits(sd, x, torch.full((x.shape[0],), sigma), emulate_bf16=True)
  m = m.to(cuda)
  got = m.backbone.packed().forward(x.to(cuda), sigma).cpu()
  scale = float(ref.abs().max())
  e_emu, e_ref = float((got - emu).abs().max()) / scale, float((got - ref).abs().max()) / scale
  print(f'\n[dit {tag}] scale={scale:.4g} rel.err vs bf16-emulating oracle={e_emu:.3e} vs fp32 reference={e_ref:.3e}')
  assert e_emu < TIGHT and e_ref < LOOSE
  got8 = m.backbone.packed().forward(x.to(cuda).to(torch.uint8), sigma).cpu()
  assert torch.equal(got8, got)
  # what the sampler sees: post-SUBS log-probs of masked positions, and the Tweedie / noise-removal argmax
  lp_got = svdd.subs_parameterization(got, x)[x == 4][:, :4]
  lp_ref = svdd.subs_parameterization(ref, x)[x == 4][:, :4]
  assert float((lp_got - lp_ref).abs().max()) < 0.08
  # reference API: Diffusion.forward(x, sigma) -> post-SUBS log-probs (diffusion_gosai.py:339-357)
  m.time_conditioning = sigma != 0.0
  lp = m.forward(x.to(cuda), torch.full((x.shape[0],), sigma, device=cuda)).cpu()
  assert float((lp[x == 4][:, :4] - lp_ref).abs().max()) < 0.08


def test_dit_batch_shapes_and_chunking(cuda):
  """Per-sequence logits do not depend on the batch composition: ragged query / key tiles
  (L = 77), several chunks of the workspace (rows > 32768), duplicates."""
  m = helpers.build_dit(n_blocks=2, length=77).to(cuda)
  den = m.backbone.packed()
  x = helpers.random_tokens(450, 77, 3, 0.6).to(cuda)
  x[300] = x[7]
  full = den.forward(x, 0.0)
  assert torch.isfinite(full).all() and torch.equal(full[300], full[7])
  for lo, n in ((0, 1), (7, 3), (425, 25)):
    part = den.forward(x[lo:lo + n].contiguous(), 0.0)
    assert torch.equal(part, full[lo:lo + n]), (lo, n)


@pytest.mark.parametrize('mode', ['plain', 'mc', 'pm'])
def test_dit_as_stage_1_every_transition_matches_oracle(cuda, mode):
  """`backbone: dit` in the SVDD loops (eager and through the CUDA graph): every transition is
  the oracle's given the kernels' logits / values and the injected noise; graph replay == eager."""
  m = helpers.build_dit(n_blocks=2, length=50).to(cuda)
  B, M, steps, L = 5, 4, 6, 50
  U = torch.rand(steps, M, B, L, 5, generator=torch.Generator().manual_seed(6))
  noise = diffusion_gosai.InjectedNoise(U.to(cuda))
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  rm = value_nets.OriBaseModel(emb, head)
  trace = []
  if mode == 'plain':
    x = m.decode_sample(num_steps=steps, eval_sp_size=B, noise=diffusion_gosai.InjectedNoise(U[:, :1].to(cuda)), trace=trace)
  elif mode == 'mc':
    x = m.controlled_sample(emb, head, num_steps=steps, eval_sp_size=B, sample_M=M, noise=noise, trace=trace)
  else:
    x = m.controlled_sample_tweedie(rm, num_steps=steps, eval_sp_size=B, sample_M=M, noise=noise, options='True',
                                    task='rna', trace=trace)
  assert x.shape == (B, L) and int(x.max()) <= 3
  sd = {k: v.cpu() for k, v in m.state_dict().items()}
  Mm = 1 if mode == 'plain' else M
  for rec in trace[:-1]:
    xs, lg = rec['x'].cpu().long(), rec['logits'].cpu()
    q = svdd.build_q_xs(svdd.subs_parameterization(lg, xs), rec['mc_t'], rec['mc_s'])
    cand = svdd.draw_candidates(xs, q, U[rec['step'], :Mm])
    if mode == 'plain':
      assert torch.equal(cand[0], rec['x_next'].cpu().long())
    else:
      assert torch.equal(cand, rec['cand'].cpu().long())
      if mode == 'pm':
        x0 = svdd.subs_parameterization(rec['logits2'].cpu(), cand.reshape(M * B, L)).argmax(-1)
        assert torch.equal(x0, rec['x0'].cpu().long())
      idx = svdd.select(rec['scores'].cpu().t().contiguous())
      assert torch.equal(svdd.gather_selected(cand, idx), rec['x_next'].cpu().long())
    with torch.no_grad():
      ref = nets.dit_logits(sd, xs)
    assert float((lg - ref).abs().max() / ref.abs().max()) < LOOSE
  # product path: in-kernel noise, eager == graph replay
  if mode != 'plain':
    run = (lambda: m.controlled_sample(emb, head, num_steps=steps, eval_sp_size=B, sample_M=M)) if mode == 'mc' else \
          (lambda: m.controlled_sample_tweedie(rm, num_steps=steps, eval_sp_size=B, sample_M=M, options='True', task='rna'))
    m.manual_seed(11)
    m.use_cuda_graph = False
    a = run()
    m.manual_seed(11)
    m.use_cuda_graph = True
    b = run()
    assert torch.equal(a, b)


def test_dit_time_conditioning_and_checkpoint_keys(cuda, tmp_path):
  """sigma reaches the kernels through the host-evaluated adaLN vectors; a Lightning checkpoint
  with the reference's ``backbone.*`` DiT keys round-trips through load_from_checkpoint."""
  m = helpers.build_dit(n_blocks=2, length=50, time_conditioning=True).to(cuda)
  x = helpers.random_tokens(3, 50, 2, 0.5).to(cuda)
  a = m.backbone.packed().forward(x, 0.2)
  b = m.backbone.packed().forward(x, 0.9)
  assert float((a - b).abs().max()) > 1e-3
  sd = {k: v.cpu() for k, v in m.state_dict().items()}
  with torch.no_grad():
    ref = nets.dit_logits(sd, x.cpu(), torch.full((3,), 0.9))
  assert float((b.cpu() - ref).abs().max() / ref.abs().max()) < LOOSE
  path = tmp_path / 'dit.ckpt'
  torch.save({'state_dict': sd, 'epoch': 1}, path)
  m2 = diffusion_gosai.Diffusion.load_from_checkpoint(str(path), config=m.config).to(cuda).eval()
  assert torch.equal(m2.backbone.packed().forward(x, 0.9), b)
  # the reverse-step API with time conditioning on (sigma_t per step)
  t = torch.full((3, 1), 0.5, device=cuda)
  xn, _, q, _ = m._ddpm_update_finetune(x, t, (1 - 1e-5) / 128)
  keep = (x != 4)
  assert torch.equal(xn[keep], x[keep]) and q.shape == (3, 50, 5)


@pytest.mark.parametrize('L,n', [(200, 9), (50, 33), (256, 2), (77, 5), (16, 3)])
def test_dit_attention_resident_kernel_matches_streaming_kernel(cuda, L, n, monkeypatch):
  """L <= 256: the whole-sequence-resident attention kernel (K / V rotated and staged once per
  (sequence, head)) against the streaming kernel it replaces (SVDD_DIT_ATTN_STREAM=1, read per
  call): same rotary, same bf16 roundings, same key order within the online softmax."""
  m = helpers.build_dit(n_blocks=2, length=L).to(cuda)
  den = m.backbone.packed()
  x = helpers.random_tokens(n, L, 17, 0.5).to(cuda)
  monkeypatch.setenv('SVDD_DIT_ATTN_STREAM', '1')
  a = den.forward(x, 0.0).clone()
  monkeypatch.setenv('SVDD_DIT_ATTN_STREAM', '0')
  b = den.forward(x, 0.0)
  err = float((a - b).abs().max()) / float(a.abs().max())
  print(f'\n[dit attention resident vs streaming L={L} n={n}] rel.err {err:.3e}')
  assert err < 2e-3
