"""Per-launch warm timing of one reverse step's network passes (value net on B*M candidates,
denoiser on B): prints the conv_gemm table (SVDD_PROF_DUMP) and the whole-pass times, eager
and as a CUDA graph.  Tuning aid; bench.py is the measurement of record.
    python tools/gemm_table.py [--B 128] [--M 10]
"""
import argparse
import os
import sys

os.environ['SVDD_PROF_DUMP'] = '1'
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import bench  # noqa: E402
from svdd_b200 import _lib, value_nets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--B', type=int, default=bench.B_PER_GPU)
ap.add_argument('--M', type=int, default=bench.M)
args = ap.parse_args()
dev = torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
B, M, L = args.B, args.M, bench.L
cand = torch.randint(0, 4, (M * B, L), device=dev, dtype=torch.uint8)
scorer = value_nets.packed_scorer(emb, head)
den = model.backbone.packed()
sc_out = torch.empty(M * B, dtype=torch.float32, device=dev)
lg_out = torch.empty(B, L, 5, dtype=torch.float32, device=dev)
for _ in range(2):
  scorer.score(cand, out=sc_out); den.forward(cand[:B], 0.0, out=lg_out)
torch.cuda.synchronize()
_lib.profile_begin()
scorer.score(cand, out=sc_out)
den.forward(cand[:B], 0.0, out=lg_out)
torch.cuda.synchronize()
ms, n, fl = _lib.profile_end()
print(f'conv_gemm: {n} launches, {ms:.3f} ms summed, {fl / ms / 1e9:.1f} TFLOP/s')


def timed(fn, reps=5):
  ts = []
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return min(ts), sorted(ts)[len(ts) // 2]


print('eager  value pass   min/median ms', timed(lambda: scorer.score(cand, out=sc_out)))
print('eager  denoiser     min/median ms', timed(lambda: den.forward(cand[:B], 0.0, out=lg_out)))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
  scorer.score(cand, out=sc_out); den.forward(cand[:B], 0.0, out=lg_out)
  torch.cuda.synchronize()
  with torch.cuda.graph(g, stream=s):
    scorer.score(cand, out=sc_out)
    den.forward(cand[:B], 0.0, out=lg_out)
torch.cuda.synchronize()
print('graph  value+denoiser min/median ms', timed(g.replay))
