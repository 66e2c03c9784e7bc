"""-m gpu: the tensor-core networks (through the C ABI) against the reference
goldens (fp32) and the oracle's bf16-operand emulation.

Stated tolerances (north_star: "network logits and values agree within a stated
bf16/fp32 tolerance"):
  * vs the oracle with bf16-rounded GEMM operands (same rounding points, fp32
    accumulate): max |d| <= TIGHT * scale  -- catches semantic errors;
  * vs the reference's fp32 output: max |d| <= LOOSE * scale -- the bf16 budget.
scale = max |reference output| over the batch."""
import numpy as np
import pytest
import torch

import helpers
from oracle import nets, svdd
from svdd_b200 import _lib, value_nets

pytestmark = pytest.mark.gpu

TIGHT, LOOSE = 4e-3, 3e-2
# value nets: random-init nets emit near-constant scores much smaller than their
# internal activations (O(1)), so the error is stated as absolute for |v| <= 1 and
# relative beyond: |d| <= tol * max(1, max|v|).
V_TIGHT, V_LOOSE = 2e-3, 1e-2


def T(a):
  return torch.from_numpy(np.asarray(a))


def _report(name, got, emu, ref, floor=0.0):
  scale = max(float(ref.abs().max()), floor)
  e_emu = float((got - emu).abs().max()) / scale
  e_ref = float((got - ref).abs().max()) / scale
  print(f'\n[{name}] scale={scale:.4g} rel.err vs bf16-emulating oracle={e_emu:.3e} vs fp32 reference={e_ref:.3e}')
  return e_emu, e_ref


@pytest.mark.parametrize('L', [50, 200])
@pytest.mark.parametrize('tok_dtype', [torch.int64, torch.uint8])
def test_denoiser_logits(cuda, L, tok_dtype):
  g = helpers.load_golden('denoiser_seed44.npz')
  m = helpers.build_denoiser(44, L)
  sd = {'backbone.' + k: v for k, v in m.state_dict().items()}
  x = T(g[f'L{L}_tokens'])
  ref = T(g[f'L{L}_logits'])
  with torch.no_grad():
    emu = nets.denoiser_logits(sd, x, emulate_bf16=True)
  m = m.to(cuda)
  got = m.packed().forward(x.to(cuda).to(tok_dtype), 0.0).cpu()
  e_emu, e_ref = _report(f'denoiser L={L}', got, emu, ref)
  assert e_emu < TIGHT and e_ref < LOOSE
  # the distribution the sampler sees: max |d log p| over real tokens of masked rows
  lp_got = svdd.subs_parameterization(got, x)[x == 4][:, :4]
  lp_ref = T(g[f'L{L}_log_p'])[x == 4][:, :4]
  assert float((lp_got - lp_ref).abs().max()) < 0.05
  # argmax tokens (noise removal / Tweedie) agree except where the top-2 gap is tiny
  a_got = _lib.x0_argmax(got.to(cuda), x.to(cuda)).cpu()
  a_ref = svdd.subs_parameterization(ref, x)[:, :, :4].argmax(-1)
  gap = svdd.subs_parameterization(ref, x)[:, :, :4].topk(2, -1).values
  bad = a_got != a_ref
  assert float((gap[..., 0] - gap[..., 1])[bad].max() if bad.any() else 0.0) < 0.05


def test_denoiser_batch_shapes(cuda):
  """Ragged batch sizes (tiles that straddle sequences / partial last tile) give
  the same per-sequence logits as a batch of one."""
  for L in (50, 200):
    m = helpers.build_denoiser(44, L).to(cuda)
    x = helpers.random_tokens(7, L, 3, 0.6).to(cuda)
    full = m.packed().forward(x, 0.0)
    for n in (1, 2, 5):
      part = m.packed().forward(x[:n].contiguous(), 0.0)
      assert torch.equal(part, full[:n]), (L, n)


@pytest.mark.parametrize('L,n', [(200, 150), (200, 3), (50, 301), (50, 1), (64, 5), (100, 9), (130, 4), (256, 2),
                                 (24, 7), (33, 601), (32, 2)])
def test_denoiser_fused_kernel_matches_layer_by_layer(cuda, L, n, monkeypatch):
  """The persistent whole-network kernel (activations in smem / TMEM, taps as descriptor row
  offsets) against the layer-by-layer conv_gemm path: same bf16 rounding points, so the two
  agree to fp32-accumulation-order noise.  Covers two row tiles (L > 128), one tile
  (64 < L <= 128), two sequences per CTA (L <= 64, odd count) and several items per CTA."""
  m = helpers.build_denoiser(44, L).to(cuda)
  x = helpers.random_tokens(n, L, 5, 0.6).to(cuda)
  den = m.packed()
  monkeypatch.setenv('SVDD_DEN_FUSED', '0')
  layered = den.forward(x, 0.3)
  monkeypatch.setenv('SVDD_DEN_FUSED', '1')
  before = _lib.launch_count()
  fused = den.forward(x, 0.3)
  torch.cuda.synchronize()
  assert _lib.launch_count() - before == 1, 'the fused path is one launch'
  torch.cuda.synchronize()
  scale = float(layered.abs().max())
  err = float((fused - layered).abs().max()) / scale
  print(f'\n[denoiser fused vs layered L={L} n={n}] rel.err {err:.3e}')
  assert err < 6e-3


@pytest.mark.parametrize('L,n', [(50, 333), (20, 9), (64, 2)])
def test_denoiser_split_epilogue_matches_one_thread_per_row(cuda, L, n, monkeypatch):
  """Short sequences (two per CTA): the split epilogue (two threads per row, LayerNorm statistics
  combined pairwise) against the one-thread-per-row epilogue (SVDD_DEN_SPLIT=0, read per call).
  Same bf16 rounding points; the statistics differ only in summation order."""
  m = helpers.build_denoiser(44, L).to(cuda)
  x = helpers.random_tokens(n, L, 11, 0.5).to(cuda).to(torch.uint8)
  den = m.packed()
  monkeypatch.setenv('SVDD_DEN_SPLIT', '0')
  one = den.forward(x, 0.0).clone()
  monkeypatch.setenv('SVDD_DEN_SPLIT', '1')
  two = den.forward(x, 0.0)
  again = den.forward(x, 0.0).clone()
  torch.cuda.synchronize()
  assert torch.equal(two, again), 'split epilogue is not deterministic'
  err = float((two - one).abs().max()) / float(one.abs().max())
  print(f'\n[denoiser split vs single epilogue L={L} n={n}] rel.err {err:.3e}')
  assert err < 6e-3


@pytest.mark.parametrize('L,n', [(50, 333), (33, 40), (64, 3), (20, 9)])
def test_denoiser_combined_planes_match_two_tile_scheme(cuda, L, n, monkeypatch):
  """Short sequences: combined operand planes (ONE 128-row MMA per tap serves both sequences of a
  CTA for taps that fit the zero gap between them, the rest on isolated planes with their own
  accumulators) against the two-tile scheme (SVDD_DEN_CMB=0, read per call), and the 16-warp quad
  epilogue (SVDD_DEN_EW=16) against the default 8-warp split epilogue.  Same bf16 rounding
  points; the sums differ in accumulation order only."""
  m = helpers.build_denoiser(44, L).to(cuda)
  x = helpers.random_tokens(n, L, 13, 0.5).to(cuda).to(torch.uint8)
  den = m.packed()
  monkeypatch.setenv('SVDD_DEN_ILV', '0')
  monkeypatch.setenv('SVDD_DEN_CMB', '0')
  two = den.forward(x, 0.0).clone()
  monkeypatch.setenv('SVDD_DEN_CMB', '1')
  cmb = den.forward(x, 0.0).clone()
  assert torch.equal(cmb, den.forward(x, 0.0)), 'combined mode is not deterministic'
  monkeypatch.setenv('SVDD_DEN_EW', '16')
  q16 = den.forward(x, 0.0).clone()
  monkeypatch.delenv('SVDD_DEN_EW')
  scale = float(two.abs().max())
  e1, e2 = float((cmb - two).abs().max()) / scale, float((q16 - two).abs().max()) / scale
  print(f'\n[denoiser combined vs two-tile L={L} n={n}] rel.err {e1:.3e}; 16-warp epilogue {e2:.3e}')
  assert e1 < 6e-3 and e2 < 6e-3
  # odd batch: the last CTA item holds one sequence only
  one = den.forward(x[:1].contiguous(), 0.0)
  assert torch.equal(one[0], cmb[0])
  # interleaved planes (row 2p / 2p+1 = position p of sequence A / B; one MMA per tap for both), one item per CTA ...
  monkeypatch.setenv('SVDD_DEN_ILV', '1')
  monkeypatch.setenv('SVDD_DEN_PAIR', '0')
  ilv1 = den.forward(x, 0.0).clone()
  # ... and the default: two items (four sequences) in flight per CTA (csrc/den_short.cuh)
  monkeypatch.setenv('SVDD_DEN_PAIR', '1')
  ilv = den.forward(x, 0.0).clone()
  assert torch.equal(ilv, ilv1), 'two items in flight must not change a single bit'
  # ... and as CTA pairs that split every weight tile (cta_group::2 MMAs; an experiment kept behind SVDD_DEN_CG=2)
  monkeypatch.setenv('SVDD_DEN_CG', '2')
  assert torch.equal(den.forward(x, 0.0), ilv), 'CTA pairs must not change a single bit'
  for k in (1, 3, 5):
    if k <= n:
      assert torch.equal(den.forward(x[:k].contiguous(), 0.0), ilv[:k]), k
  monkeypatch.delenv('SVDD_DEN_CG')
  for k in (1, 2, 3, 5):          # partial pairs / items at the end of the batch
    if k <= n:
      assert torch.equal(den.forward(x[:k].contiguous(), 0.0), ilv[:k]), k
  assert torch.equal(ilv, den.forward(x, 0.0)), 'interleaved mode is not deterministic'
  e3 = float((ilv - two).abs().max()) / scale
  print(f'[denoiser interleaved vs two-tile L={L} n={n}] rel.err {e3:.3e}')
  assert e3 < 6e-3
  assert torch.equal(den.forward(x[:1].contiguous(), 0.0)[0], ilv[0])
  if n >= 3:      # a sequence gives the same logits in the A slot and in the B slot of an item
    assert torch.equal(den.forward(x[1:3].contiguous(), 0.0)[0], ilv[1])


def test_convgru_value(cuda):
  g = helpers.load_golden('value_nets.npz')
  tok = T(g['convgru_tokens'])
  oh = svdd.transform_samples(tok).float()
  for name, build, key, tr in (('value', helpers.build_convgru_value, 'convgru_values', False),
                               ('oracle', helpers.build_convgru_oracle, 'rnaoracle_values', True)):
    emb, head = build()
    inp = oh.transpose(1, 2) if tr else oh
    with torch.no_grad():
      emu = nets.convgru_value(emb.state_dict(), head.state_dict(), inp, emulate_bf16=True).reshape(-1)
    ref = T(g[key])
    emb, head = emb.to(cuda), head.to(cuda)
    got = value_nets.score_tokens(emb, head, tok.to(cuda)).cpu()
    e_emu, e_ref = _report(f'convgru {name}', got, emu, ref, floor=1.0)
    assert e_emu < V_TIGHT and e_ref < V_LOOSE
    got8 = value_nets.score_tokens(emb, head, tok.to(cuda).to(torch.uint8)).cpu()
    assert torch.equal(got8, got)


def test_convgru_many_rows(cuda):
  """More rows than one chunk / one GRU block: per-row results do not depend on
  batch composition."""
  emb, head = helpers.build_convgru_value()
  emb, head = emb.to(cuda), head.to(cuda)
  tok = helpers.random_tokens(1000, 50, 21, 0.5).to(cuda)
  full = value_nets.score_tokens(emb, head, tok)
  part = value_nets.score_tokens(emb, head, tok[137:150].contiguous())
  assert torch.allclose(part, full[137:150], rtol=0, atol=1e-6)
  assert torch.isfinite(full).all()


def test_convgru_tensor_core_recurrence_equals_scalar_kernel(cuda, tmp_path):
  """The mma.sync GRU recurrence (bf16 hi/lo split, fp32 accumulate) against the scalar
  FMA-pipe kernel (SVDD_GRU_SCALAR=1, read once per process: run in a subprocess) on rows
  that do not fill the last 16-sequence block."""
  import os
  import subprocess
  import sys
  out = str(tmp_path / 'scalar.pt')
  code = ('import sys, torch; sys.path[:0] = [%r, %r]; import helpers; from svdd_b200 import value_nets; '
          'd = torch.device("cuda:0"); e, h = helpers.build_convgru_value(); '
          'tok = helpers.random_tokens(77, 50, 23, 0.4).to(d); '
          'torch.save(value_nets.score_tokens(e.to(d), h.to(d), tok).cpu(), %r)'
          % (helpers.ROOT, os.path.join(helpers.ROOT, 'tests'), out))
  subprocess.run([sys.executable, '-c', code], check=True, env=dict(os.environ, SVDD_GRU_SCALAR='1'), timeout=300)
  scalar = torch.load(out)
  emb, head = helpers.build_convgru_value()
  tok = helpers.random_tokens(77, 50, 23, 0.4).to(cuda)
  got = value_nets.score_tokens(emb.to(cuda), head.to(cuda), tok).cpu()
  err = float((got - scalar).abs().max())
  print(f'\n[convgru mma vs scalar recurrence] max |d| = {err:.3e} (scale {float(scalar.abs().max()):.3g})')
  assert err <= 2e-5 * max(1.0, float(scalar.abs().max()))


def test_enformer_value(cuda):
  g = helpers.load_golden('value_nets.npz')
  tok = T(g['enformer_tokens'])
  oh = svdd.transform_samples(tok).float()
  emb, head = helpers.build_enformer()
  with torch.no_grad():
    emu = nets.enformer_value(emb.state_dict(), head.state_dict(), oh, n_heads=8,
                              emulate_bf16=True).reshape(-1)
  ref = T(g['enformer_values'])
  emb, head = emb.to(cuda), head.to(cuda)
  got = value_nets.score_tokens(emb, head, tok.to(cuda)).cpu()
  print('\n got', got.numpy(), '\n emu', emu.numpy(), '\n ref', ref.numpy())
  # The trunk is ~45 bf16-operand layers deep; with calibrated BatchNorm the rounding
  # noise reaches ~0.7% of the activation scale at the trunk output (measured stage by
  # stage with tools/debug_enformer.py: the kernels track the fp32 reference slightly
  # BETTER than the CPU bf16 emulation does, and no stage shows a jump).  Stated
  # tolerance: 10% of the score spread over the batch, against both targets.
  spread = float(ref.max() - ref.min())
  e_emu, e_ref = _report('enformer (384ch, 2 blocks)', got, emu, ref, floor=spread)
  assert e_emu < 0.1 and e_ref < 0.1
  assert torch.equal(got.argsort(), ref.argsort())
  got8 = value_nets.score_tokens(emb, head, tok.to(cuda).to(torch.uint8)).cpu()
  assert torch.equal(got8, got)


# Full headline network (decode.py:78-80): stated tolerances, ABSOLUTE on the score.  Scores of
# this seeded net have mean 0.16, std 0.11, range 0.58 over the batch.  The bf16-operand pipeline
# (7 conv stages + 11 transformer blocks, ~50 GEMMs deep, fp32 accumulate / normalisation /
# residual stream) lands within EF_REF of the REFERENCE's fp32 output (measured 1.9e-2 max,
# 3.5e-3 mean = 3 % of the std).  The oracle's CPU bf16-operand emulation is a second bf16
# pipeline with its own rounding noise of the same size (emulation vs fp32: 2.2e-2), so kernel vs
# emulation is bounded by the SUM (EF_EMU; measured 3.7e-2) -- at this depth it documents the
# noise level, the fp32 bound is the parity statement.  The 384-channel / 2-block net above
# (test_enformer_value) is where the emulation is tight enough to catch semantic errors.
EF_REF, EF_MEAN, EF_EMU = 2.5e-2, 5e-3, 4.5e-2


def test_enformer_full_headline_network_parity(cuda):
  """A10 on the network the headline number is measured on: EnformerTrunk(7, 1536, 11, 8, 64) +
  ConvHead(1, 3072) (decode.py:78-80), calibrated BatchNorm, perturbed to_out; 32 sequences x
  M = 10 SVDD-step-shaped candidates at L = 200, against (1) the reference's own fp32 outputs
  (tests/golden/enformer_full.npz) and (2) the oracle's bf16-operand emulation.  What selection
  consumes is the per-sequence argmax over the M candidates (diffusion_gosai.py:1219-1227):
  agreement with the fp32 reference and the regret of disagreements are reported and bounded."""
  g = helpers.load_golden('enformer_full.npz')
  S, M, L = 32, 10, 200
  x, cand = helpers.svdd_step_candidates(S, M, L, seed=2024)
  assert np.array_equal(cand.numpy(), g['cand'])
  ref = T(g['values'])                                        # [M, S] fp32, reference modules
  emb, head = helpers.build_enformer(full=True)
  np.testing.assert_array_equal(helpers.state_checksum(emb.state_dict()), g['checksum'])
  with torch.no_grad():
    emu = nets.enformer_value(emb.state_dict(), head.state_dict(),
                              svdd.transform_samples(cand.reshape(M * S, L)).float(), n_heads=8,
                              emulate_bf16=True).reshape(M, S)
  emb, head = emb.to(cuda), head.to(cuda)
  got = value_nets.score_tokens(emb, head, cand.reshape(M * S, L).to(cuda)).cpu().reshape(M, S)
  sd = float(ref.std())
  e_emu, e_ref = float((got - emu).abs().max()), float((got - ref).abs().max())
  m_ref = float((got - ref).abs().mean())
  print(f'\n[enformer FULL 1536ch/11 blocks, {M * S} candidates] score mean {float(ref.mean()):.4f} std {sd:.4f} '
        f'range {float(ref.max() - ref.min()):.4f}\n  max|d| vs bf16-emulating oracle {e_emu:.3e} ({e_emu / sd:.1%} of std), '
        f'vs fp32 reference {e_ref:.3e} ({e_ref / sd:.1%} of std), mean|d| vs fp32 {m_ref:.3e}; '
        f'(emulation vs fp32: {float((emu - ref).abs().max()):.3e})')
  # correlation with the reference over all candidates
  corr = float(torch.corrcoef(torch.stack([got.reshape(-1), ref.reshape(-1)]))[0, 1])
  # selection: argmax over each sequence's M candidates
  spread = ref.max(0).values - ref.min(0).values
  pick, pick_ref, pick_emu = got.argmax(0), ref.argmax(0), emu.argmax(0)
  regret = ref.max(0).values - ref.gather(0, pick[None])[0]
  agree = float((pick == pick_ref).float().mean())
  print(f'  argmax over M=10: agreement with fp32 reference {agree:.3f} (the CPU bf16 emulation: '
        f'{float((pick_emu == pick_ref).float().mean()):.3f}); regret of disagreements max {float(regret.max()):.3e} '
        f'mean {float(regret.mean()):.3e}; mean spread over M {float(spread.mean()):.3e}')
  print(f'  correlation with the fp32 reference over the {M * S} scores: {corr:.5f}')
  assert e_ref <= EF_REF and m_ref <= EF_MEAN and e_emu <= EF_EMU
  assert corr >= 0.995
  assert agree >= 0.85 and float(regret.max()) <= 0.25 * float(spread.mean())
  assert float(regret.mean()) <= 0.03 * float(spread.mean())
  # uint8 tokens take the same path
  got8 = value_nets.score_tokens(emb, head, cand.reshape(M * S, L).to(cuda).to(torch.uint8)).cpu().reshape(M, S)
  assert torch.equal(got8, got)


def test_enformer_identical_candidates_score_identically(cuda):
  """Late in a trajectory most of a sequence's M candidates are IDENTICAL token rows; the
  reference then sees exact ties and argmax takes the first (diffusion_gosai.py:1224).  The
  one-call scoring of all M*B rows must give bit-identical scores to identical rows wherever
  they sit in the batch, or the tie-break would depend on tile placement."""
  emb, head = helpers.build_enformer(full=True)
  emb, head = emb.to(cuda), head.to(cuda)
  base = helpers.random_tokens(37, 200, 91, 0.1)
  rows = torch.cat([base, base[:5], base.flip(0), base[3:4].expand(9, -1)], 0)          # 88 rows, duplicates far apart
  got = value_nets.score_tokens(emb, head, rows.to(cuda)).cpu()
  assert torch.equal(got[:37], got[42:79].flip(0)) and torch.equal(got[:5], got[37:42])
  assert bool((got[79:] == got[3]).all())
  S, M = 16, 10
  x, cand = helpers.svdd_step_candidates(S, M, 200, seed=7, p_masks=(0.03,))
  sc = value_nets.score_tokens(emb, head, cand.reshape(M * S, 200).to(cuda)).cpu().reshape(M, S)
  same = (cand == cand[:1]).all(-1)                                                   # [M, S]: identical to candidate 0
  assert int(same.sum()) > S                                                          # the case occurs
  assert torch.equal(sc[same], sc[:1].expand(M, S)[same])


def test_dna_reward_oracle_scores_task_0_of_3(cuda):
  """A12: the DNA reward oracle has 3 tasks (oracle.py:72) and the path scores
  ``reward_model(onehot.transpose(1, 2))[:, 0]`` (diffusion_gosai.py:1430).  ConvHead(n_tasks=3)
  through the CUDA scorer == task 0 of the reference's [N, 3] output, within the value-net
  tolerance, and NOT tasks 1 / 2."""
  from svdd_b200 import synthetic
  g = helpers.load_golden('enformer_full.npz')
  _, cand = helpers.svdd_step_candidates(32, 10, 200, seed=2024)
  rm = synthetic.build_dna_reward_model()
  np.testing.assert_array_equal(helpers.state_checksum(rm.head.state_dict()), g['head3_checksum'])
  rm = rm.to(cuda)
  tok = cand[:2].reshape(64, 200).to(cuda)
  got = value_nets.score_tokens(rm.embedding, rm.head, tok).cpu().reshape(2, 32)
  ref3 = T(g['values3'])                                                               # [2, 32, 3]
  e = [float((got - ref3[..., t]).abs().max()) for t in range(3)]
  print(f'\n[dna reward oracle] max|d| vs task 0 / 1 / 2 of the reference: {e[0]:.3e} / {e[1]:.3e} / {e[2]:.3e}')
  assert e[0] <= EF_REF and min(e[1], e[2]) > 3 * EF_REF


def test_enformer_batch_independence(cuda):
  emb, head = helpers.build_enformer()
  emb, head = emb.to(cuda), head.to(cuda)
  tok = helpers.random_tokens(70, 200, 33, 0.5).to(cuda)
  full = value_nets.score_tokens(emb, head, tok)
  part = value_nets.score_tokens(emb, head, tok[11:14].contiguous())
  assert torch.allclose(part, full[11:14], rtol=0, atol=1e-5)
  assert torch.isfinite(full).all()


@pytest.mark.parametrize('full,n_cand', [(False, 3), (False, 70), (False, 300), (True, 128), (True, 1280)])
def test_enformer_persistent_tower_matches_per_launch_path(cuda, full, n_cand, monkeypatch):
  """The one-launch transformer tower (csrc/tower.cuh: flag-synchronised work items over row
  tiles) against the launch-per-GEMM path it replaces: same GEMM k order, same LayerNorm and
  attention arithmetic, so the scores agree to fp32 rounding of the final BN+GELU operand.
  Row counts cover a single ragged row tile, several tiles, and the c2 size (2560 rows); the
  small net (384 channels) has ragged column tiles, the full one (1536) is the bench network."""
  emb, head = helpers.build_enformer(full=full)
  emb, head = emb.to(cuda), head.to(cuda)
  tok = helpers.random_tokens(n_cand, 200, 7 + n_cand, 0.5).to(cuda)
  monkeypatch.setenv('SVDD_TOWER', '0')
  ref = value_nets.score_tokens(emb, head, tok).cpu()
  monkeypatch.setenv('SVDD_TOWER', '1')
  monkeypatch.setenv('SVDD_TOWER_SPLITK', '1')       # one K slice per tile: the per-launch path's summation order
  before = _lib.launch_count()
  got = value_nets.score_tokens(emb, head, tok).cpu()
  launches = _lib.launch_count() - before
  again = [value_nets.score_tokens(emb, head, tok).cpu() for _ in range(4)]
  scale = float(ref.abs().max())
  err = float((got - ref).abs().max())
  print(f'\n[tower full={full} n_cand={n_cand}] launches={launches} max|d|={err:.3e} scale={scale:.3e}')
  assert torch.isfinite(got).all()
  for k, o in enumerate(again):
    assert torch.equal(got, o), f'persistent tower is not deterministic (run {k + 1}: {int((o != got).sum())} scores differ, ' \
                                f'max {float((o - got).abs().max()):.3e})'
  assert err <= 2e-3 * max(scale, 1e-3)
  # optional: FF2's K split in two slices whose partial sums are added IN ORDER -- the fp32 residual
  # stream differs from the single-sum path in the last bit, which bf16 roundings downstream amplify
  # to the level of the documented bf16 noise (<= 10 % of the scale); still run-to-run identical
  monkeypatch.setenv('SVDD_TOWER_SPLITK', '2')
  sk = [value_nets.score_tokens(emb, head, tok).cpu() for _ in range(4)]
  err_sk = float((sk[0] - ref).abs().max())
  print(f'[tower split-K] max|d|={err_sk:.3e}')
  for o in sk[1:]:
    assert torch.equal(sk[0], o), 'ordered split-K is not deterministic'
  assert err_sk <= 0.1 * max(scale, 1e-3)


def test_value_rank_agreement(cuda):
  """What selection needs: the candidate the fp32 oracle prefers is (nearly always)
  the one the bf16 tensor-core net prefers.  Reported, with a weak bound."""
  emb, head = helpers.build_convgru_value()
  x = helpers.random_tokens(64, 50, 5, 0.5)
  cand = torch.stack([torch.where(x == 4, torch.randint(0, 5, x.shape,
                      generator=torch.Generator().manual_seed(m)), x) for m in range(10)])
  with torch.no_grad():
    ref = nets.convgru_value(emb.state_dict(), head.state_dict(),
                             svdd.transform_samples(cand.reshape(-1, 50)).float()).reshape(10, 64)
  got = value_nets.score_tokens(emb.to(cuda), head.to(cuda), cand.reshape(-1, 50).to(cuda)).cpu().reshape(10, 64)
  agree = float((got.argmax(0) == ref.argmax(0)).float().mean())
  regret = float((ref.max(0).values - ref.gather(0, got.argmax(0)[None])[0]).max())
  spread = float((ref.max(0).values - ref.min(0).values).mean())
  print(f'\n[rank] argmax agreement={agree:.3f} worst regret={regret:.2e} mean spread={spread:.2e}')
  assert agree >= 0.8 and regret <= 0.25 * spread


@pytest.mark.parametrize('full,n_cand', [(False, 70), (True, 130)])
def test_enformer_pool16_matches_slab_variant(cuda, full, n_cand, monkeypatch):
  """The residual 1x1 conv (EPI_PAIR) and difference pooling (EPI_POOL2) with 16 epilogue warps
  (thread = row x 16 columns of a 32-column slab; the default) against the 8-warp kernels on the
  same slabs (SVDD_PAIR16=0 / SVDD_POOL16=0, read per call): the same arithmetic on the same bf16
  values, so the scores are bit-identical.  The small net has ragged 256-wide column tiles
  (N = 384), the full one is the bench network."""
  emb, head = helpers.build_enformer(full=full)
  emb, head = emb.to(cuda), head.to(cuda)
  tok = helpers.random_tokens(n_cand, 200, 71, 0.5).to(cuda)
  got = {}
  for pair16, pool16 in (('0', '0'), ('1', '1'), ('0', '1'), ('1', '0')):
    monkeypatch.setenv('SVDD_PAIR16', pair16)
    monkeypatch.setenv('SVDD_POOL16', pool16)
    got[pair16 + pool16] = value_nets.score_tokens(emb, head, tok).cpu()
  for k, v in got.items():
    assert torch.equal(v, got['00']), k


@pytest.mark.parametrize('full,n_cand', [(False, 70), (True, 130)])
def test_enformer_slab32_matches_wide_slabs(cuda, full, n_cand, monkeypatch):
  """EPI_PAIR / EPI_POOL2 with four 32-column slabs per column half (the default: step k + 1's TMA
  loads are in flight during step k) against the two 64-column slabs per half of rounds 1-2
  (SVDD_SLAB32=0, read per call): same arithmetic per element, bit-identical scores.  Also at a
  batch large enough for every CTA pair to walk several tiles (buffer rotation wraps)."""
  emb, head = helpers.build_enformer(full=full)
  emb, head = emb.to(cuda), head.to(cuda)
  for n in (n_cand, 3, 1):
    tok = helpers.random_tokens(n, 200, 72 + n, 0.5).to(cuda)
    monkeypatch.setenv('SVDD_SLAB32', '0')
    ref = value_nets.score_tokens(emb, head, tok).cpu()
    monkeypatch.setenv('SVDD_SLAB32', '1')
    for w16 in ('1', '0'):                 # 16-warp (default) and 8-warp kernels on the 32-column slabs
      monkeypatch.setenv('SVDD_PAIR16', w16)
      monkeypatch.setenv('SVDD_POOL16', w16)
      got = value_nets.score_tokens(emb, head, tok).cpu()
      print(f'\n[slab32 full={full} n={n} w16={w16}] max|d| = {float((got - ref).abs().max()):.3e}')
      assert torch.equal(got, ref)


def test_enformer_stem_epilogue_variants_bitwise(cuda, monkeypatch):
  """The stem GEMM (EPI_GENERIC, one staged bf16 output, K = 64: pure epilogue) on 16 warps with one
  64-column slab per column quarter (default), on 16 warps over the 32-column slabs (SVDD_EPI16=3,
  measured slower) and on the 8-warp kernel (SVDD_EPI16=0, read per call): bit-identical scores on
  the bench network, with enough candidates for every CTA pair to walk several tiles."""
  emb, head = helpers.build_enformer(full=True)
  emb, head = emb.to(cuda), head.to(cuda)
  tok = helpers.random_tokens(150, 200, 9, 0.5).to(cuda)
  got = {}
  for mode in ('0', '1', '3'):
    monkeypatch.setenv('SVDD_EPI16', mode)
    got[mode] = value_nets.score_tokens(emb, head, tok).cpu()
  assert torch.equal(got['1'], got['0']) and torch.equal(got['3'], got['0'])
  # "snake" tile order of the conv tower (every other GEMM walks its row tiles backwards, SVDD_SNAKE,
  # read per call): tiles are independent, so the order cannot change a bit
  monkeypatch.delenv('SVDD_EPI16')
  monkeypatch.setenv('SVDD_SNAKE', '0')
  fwd = value_nets.score_tokens(emb, head, tok).cpu()
  monkeypatch.setenv('SVDD_SNAKE', '1')
  assert torch.equal(value_nets.score_tokens(emb, head, tok).cpu(), fwd)
  assert torch.equal(fwd, got['1'])
  # the tower's row items drop dead intermediates from L2 (discard.global.L2, SVDD_TOWER_DISCARD, read
  # per call): a discarded line that was still needed would show up here and in the SVDD_TOWER=0 checks
  monkeypatch.setenv('SVDD_TOWER_DISCARD', '0')
  keep = value_nets.score_tokens(emb, head, tok).cpu()
  monkeypatch.setenv('SVDD_TOWER_DISCARD', '1')
  for _ in range(3):
    assert torch.equal(value_nets.score_tokens(emb, head, tok).cpu(), keep)
  assert torch.equal(keep, fwd)


@pytest.mark.parametrize('S,L,C', [(5, 100, 768), (9, 13, 128), (64, 50, 896), (1300, 2, 128), (700, 200, 256)])
def test_pair_pool_slab32_bitwise(cuda, S, L, C, monkeypatch):
  """The pair-split 1x1 conv + difference pooling in isolation, slab32 against the wide slabs:
  every output tensor (y0, yd, pooled fp32, pooled activation) bit-identical, incl. ragged column
  tiles, odd lengths and many tiles per CTA pair."""
  g = torch.Generator().manual_seed(S + L + C)
  A = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
  res = torch.randn(S, L, C, generator=g).to(cuda).bfloat16()
  W1 = (torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
  bias = torch.randn(C, generator=g).to(cuda)
  Wp = (2 * torch.eye(C) + torch.randn(C, C, generator=g) * C ** -0.5).to(cuda).bfloat16()
  scale2 = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda)
  shift2 = (0.3 * torch.randn(C, generator=g)).to(cuda)
  outs = {}
  for mode, w16 in (('0', '0'), ('1', '0'), ('1', '1')):
    monkeypatch.setenv('SVDD_SLAB32', mode)
    monkeypatch.setenv('SVDD_PAIR16', w16)
    monkeypatch.setenv('SVDD_POOL16', w16)
    outs[mode + w16] = [t.clone() for t in _lib.selftest_pair_pool(A, W1, bias, res, Wp, scale2, shift2, want_act=True)]
    torch.cuda.synchronize()
  for k in ('10', '11'):
    for a, b, name in zip(outs['00'], outs[k], ('y0', 'yd', 'pooled', 'pooled_act')):
      assert torch.equal(a, b), (k, name)


def test_module_call_surface_head_of_embedding(cuda):
  """Drop-in surface: the reference spells scoring as ``head(embedding(onehot))``
  (diffusion_gosai.py:1208-1209) and ``reward_model(onehot.transpose(1, 2))[:, 0]`` (:1430).  The
  parameter containers serve both spellings (trunk + head run as one fused pass behind a deferred
  trunk output) and return the reference's shapes; a 3-task head returns all three tasks."""
  from svdd_b200 import synthetic
  g = helpers.load_golden('enformer_full.npz')
  _, cand = helpers.svdd_step_candidates(32, 10, 200, seed=2024)
  tok = cand[0].to(cuda)
  onehot = svdd.transform_samples(cand[0]).float().to(cuda)              # [N, L, 4], masked rows all zero
  emb, head = helpers.build_enformer(full=True)
  emb, head = emb.to(cuda), head.to(cuda)
  direct = value_nets.score_tokens(emb, head, tok)
  out = head(emb(onehot))
  assert out.shape == (32, 1, 1) and torch.equal(out.reshape(-1), direct)
  rm = synthetic.build_dna_reward_model().to(cuda)
  out3 = rm(onehot.transpose(1, 2))
  assert out3.shape == (32, 3, 1)
  ref3 = T(g['values3'])[0]                                              # [32, 3] from the reference modules
  err = float((out3.squeeze(-1).cpu() - ref3).abs().max())
  print(f'\n[module call] 3-task oracle via OriBaseModel.forward: max|d| vs reference {err:.3e}')
  assert err <= EF_REF
  assert torch.equal(out3[:, 0, 0], value_nets.score_tokens(rm.embedding, rm.head, tok))
  # RNA: ConvGRU value net, both input layouts
  e2, h2 = helpers.build_convgru_value()
  e2, h2 = e2.to(cuda), h2.to(cuda)
  t2 = helpers.random_tokens(9, 50, 4, 0.4).to(cuda)
  oh2 = svdd.transform_samples(t2.cpu()).float().to(cuda)
  want = value_nets.score_tokens(e2, h2, t2)
  assert torch.equal(h2(e2(oh2)).reshape(-1), want) and torch.equal(h2(e2(oh2.transpose(1, 2))).reshape(-1), want)
  with pytest.raises(ValueError):
    head(emb(onehot * 0.5))                                              # soft inputs are not on the path


@pytest.mark.parametrize('n,L', [(77, 50), (1000, 50), (129, 50), (300, 33), (5, 64), (256, 50)])
def test_convgru_umma_recurrence_matches_mma_sync_path(cuda, n, L, monkeypatch):
  """The tcgen05 GRU recurrence with the input projection fused in (csrc/gru_umma.cuh: 256
  sequences per CTA in two groups that alternate between tensor core and gate arithmetic)
  against the mma.sync recurrence fed by the separate projection GEMM (SVDD_GRU_UMMA=0, read per
  call).  Same bf16 operands and hi / lo split; the sums differ in association only.  Row counts
  cover a partial first group, a dead second group and several CTAs."""
  for build in (helpers.build_convgru_value, helpers.build_convgru_oracle):
    emb, head = build()
    emb, head = emb.to(cuda), head.to(cuda)
    tok = helpers.random_tokens(n, L, 29 + n, 0.4).to(cuda)
    monkeypatch.setenv('SVDD_GRU_UMMA', '0')
    ref = value_nets.score_tokens(emb, head, tok).cpu()
    monkeypatch.setenv('SVDD_GRU_UMMA', '1')
    got = value_nets.score_tokens(emb, head, tok).cpu()
    again = value_nets.score_tokens(emb, head, tok).cpu()
    err = float((got - ref).abs().max())
    print(f'\n[convgru umma vs mma.sync n={n} L={L}] max |d| = {err:.3e} (scale {float(ref.abs().max()):.3g})')
    assert torch.equal(got, again) and torch.isfinite(got).all()
    assert err <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('n,L', [(77, 50), (1000, 50), (8, 50), (9, 33), (301, 62), (16, 20)])
def test_convgru_fused_conv_stack_matches_per_layer_launches(cuda, n, L, monkeypatch):
  """The persistent conv-stack kernel (csrc/cg_fused.cuh: stem + conv blocks of 8 sequences at a time
  resident in shared memory / TMEM, two sequences per 128-row tile) against the launch-per-layer path
  (SVDD_CG_FUSED=0, read per call), for the value net (BatchNorm + residual) and the oracle variant
  (bias only).  Same bf16 rounding points (every layer's output is rounded to bf16 in both), same
  fp32 accumulation over (tap, channel) -- the scores agree to accumulation-order noise.  Row counts
  cover partial items, a single item and several items per CTA; L = 62 is the longest sequence
  the two-per-tile layout admits for k5 convs."""
  for build in (helpers.build_convgru_value, helpers.build_convgru_oracle):
    emb, head = build()
    emb, head = emb.to(cuda), head.to(cuda)
    tok = helpers.random_tokens(n, L, 41 + n, 0.4).to(cuda)
    monkeypatch.setenv('SVDD_CG_FUSED', '0')
    ref = value_nets.score_tokens(emb, head, tok).cpu()
    monkeypatch.setenv('SVDD_CG_FUSED', '1')
    before = _lib.launch_count()
    got = value_nets.score_tokens(emb, head, tok).cpu()
    launches = _lib.launch_count() - before
    again = value_nets.score_tokens(emb, head, tok.to(torch.uint8)).cpu()
    err = float((got - ref).abs().max())
    print(f'\n[convgru fused conv stack n={n} L={L}] launches={launches} max |d| = {err:.3e} (scale {float(ref.abs().max()):.3g})')
    assert torch.equal(got, again) and torch.isfinite(got).all()
    assert err <= 2e-4 * max(1.0, float(ref.abs().max()))
