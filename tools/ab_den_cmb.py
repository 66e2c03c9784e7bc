"""A/B of the fused denoiser's combined mode (SVDD_DEN_CMB, read per call) on short sequences:
python tools/ab_den_cmb.py [n_rows] [L].  Prints ms per pass for both settings and max |d|."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 51200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device('cuda:0')
m = synthetic.build_denoiser(44, L).to(dev)
den = m.packed()
x = synthetic.random_tokens(n, L, 5, 0.5).to(dev).to(torch.uint8)
out = {}
for cmb in ('0', 'ilv', 'pair1', 'pair', 'pair1', 'pair'):
  os.environ['SVDD_DEN_ILV'] = '1' if cmb in ('ilv', 'pair', 'pair1') else '0'
  os.environ['SVDD_DEN_PAIR'] = '1' if cmb in ('pair', 'pair1') else '0'
  os.environ['SVDD_DEN_CG'] = '2' if cmb == 'pair' else '1'
  os.environ['SVDD_DEN_CMB'] = '0' if cmb == '0' else '1'
  y = den.forward(x, 0.0)
  torch.cuda.synchronize()
  ts = []
  for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); y = den.forward(x, 0.0); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  out[cmb] = y.clone()
  print(f'mode {cmb} (0 = two tiles, ilv = interleaved planes, pair1 = + two items in flight (default), pair = + CTA pairs (SVDD_DEN_CG=2)): n={n} L={L} min {min(ts):.3f} ms median {sorted(ts)[2]:.3f} ms')
for k in ('ilv', 'pair1', 'pair'):
  d = float((out['0'] - out[k]).abs().max())
  print(f'max |d| mode 0 vs {k}: {d:.3e} (scale {float(out["0"].abs().max()):.3g})')
