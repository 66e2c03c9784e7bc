"""CUDA-event timing of the DiT denoiser forward (configs_gosai/model/small.yaml: 12 blocks, hidden
768, 12 heads) and of its pieces.  python tools/time_dit.py [n_seq] [L]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import _lib, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device('cuda:0')
m = synthetic.build_dit(n_blocks=12, length=L).to(dev)
den = m.backbone.packed()
x = synthetic.random_tokens(n, L, 5, 0.5).to(dev).to(torch.uint8)
out = torch.empty((n, L, 5), device=dev)
for _ in range(2):
  den.forward(x, 0.0, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(5):
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); den.forward(x, 0.0, out=out); b.record()
  torch.cuda.synchronize()
  ts.append(a.elapsed_time(b))
H, F, nb = 768, 3072, 12
flops_lin = 2.0 * n * L * nb * (H * 3 * H + H * H + 2 * H * F)
flops_att = 4.0 * n * L * L * H * nb
ms = sorted(ts)[2]
print(f'DiT forward n={n} L={L}: {ms:.3f} ms  linears {flops_lin / 1e12:.2f} TFLOP + attention {flops_att / 1e12:.2f} TFLOP '
      f'-> {(flops_lin + flops_att) / ms / 1e9:.0f} TFLOP/s')
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
  den.forward(x, 0.0, out=out)
g.replay(); torch.cuda.synchronize()
ts = []
for _ in range(5):
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record(); g.replay(); b.record()
  torch.cuda.synchronize()
  ts.append(a.elapsed_time(b))
gm = sorted(ts)[2]
print(f'  as a CUDA graph: {gm:.3f} ms -> {(flops_lin + flops_att) / gm / 1e9:.0f} TFLOP/s')
if os.environ.get('SVDD_TIME_DIT_NO_PROFILE'):
  sys.exit(0)
_lib.profile_begin()
den.forward(x, 0.0, out=out)
torch.cuda.synchronize()
gms, gn, gfl = _lib.profile_end()
print(f'  tcgen05 GEMM launches: {gn}, {gms:.3f} ms summed, {gfl / gms / 1e9:.0f} TFLOP/s; everything else (LN, attention, embed, final): {ms - gms:.3f} ms')
