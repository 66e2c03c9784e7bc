"""Per-stage CUDA-event timing of one c5-shard reverse step (RNA SVDD-PM, L=50, B=1024, M=50)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers  # noqa: E402
from svdd_b200 import _lib, config, diffusion_gosai, value_nets  # noqa: E402

dev = torch.device('cuda:0')
B, M, L = 1024, 50, 50
torch.manual_seed(44)
model = diffusion_gosai.Diffusion(config.load_config('rna')).to(dev).eval()
emb, head = helpers.build_convgru_oracle()
scorer = value_nets.packed_scorer(emb.to(dev), head.to(dev))
den = model.backbone.packed()
x = torch.full((B, L), 4, dtype=torch.uint8, device=dev)
x[:, ::2] = 1
logits = torch.empty((B, L, 5), device=dev)
cand = torch.empty((M, B, L), dtype=torch.uint8, device=dev)
logits2 = torch.empty((M * B, L, 5), device=dev)
x0 = torch.empty((M * B, L), dtype=torch.uint8, device=dev)
scores = torch.empty((M, B), device=dev)
x2 = torch.empty_like(x)
stages = [
    ('denoiser B', lambda: den.forward(x, 0.0, out=logits)),
    ('subs_sample', lambda: _lib.subs_sample(logits, x, M, 0.5, 0.49, step=3, seed=1, out=cand)),
    ('denoiser B*M', lambda: den.forward(cand.reshape(M * B, L), 0.0, out=logits2)),
    ('x0_argmax', lambda: _lib.x0_argmax(logits2, cand.reshape(M * B, L), out=x0)),
    ('convgru score B*M', lambda: scorer.score(x0, out=scores.reshape(-1))),
    ('select_gather', lambda: _lib.select_gather(scores, cand, alpha=0.0, step=3, seed=1, out=x2)),
]
for name, fn in stages:
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  print(f'{name:22s} {sorted(ts)[2]:8.3f} ms')
