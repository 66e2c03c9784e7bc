"""Timeline of one pair (two items = four sequences) of den_short_kernel on CTA 0: SM clock stamps at
the hand-over points between the MMA-issuing thread and the epilogue warps (SVDD_DEN_TRACE).
Tuning aid; prints, per round, where an item's time goes.
    python tools/den_trace.py [n_rows] [L] [sm_mhz]
"""
import os
import statistics as st
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 51200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 50
mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1900.0
out = os.path.join(ROOT, 'gpurun_out', 'den_trace.csv')
os.makedirs(os.path.dirname(out), exist_ok=True)
dev = torch.device('cuda:0')
den = synthetic.build_denoiser(44, L).to(dev).packed()
x = synthetic.random_tokens(n, L, 5, 0.5).to(dev).to(torch.uint8)
os.environ['SVDD_DEN_PAIR'] = '1'
for _ in range(3):
  den.forward(x, 0.0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); den.forward(x, 0.0); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
os.environ['SVDD_DEN_TRACE'] = out
den.forward(x, 0.0)
torch.cuda.synchronize()
del os.environ['SVDD_DEN_TRACE']
rows = [[int(v) for v in l.split(',')] for l in open(out) if not l.startswith('#')]
rows = [r for r in rows if r[2] and r[9]]          # conv rounds (the final round has no epilogue stamps)
t0 = min(r[2] for r in rows)
us = lambda c: c / mhz
print(f'pass {ms:.3f} ms; clock assumed {mhz:.0f} MHz; times in us relative to the first stamp')
print('round item | mma: ready first-issue issued | epi: wait tfull ld written arrived')
for r in rows:
  print(f'{r[0]:5d} {r[1]:4d} | ' + ' '.join(f'{us(v - t0):8.2f}' for v in r[2:5]) + ' | ' + ' '.join(f'{us(v - t0):8.2f}' for v in r[5:10]) + (f' | w {us(r[10]):5.2f}' if len(r) > 10 else ''))
body = [r for r in rows if 2 <= r[0] <= 18]
def mean(f):
  return st.mean(us(f(r)) for r in body)
print('means over rounds 2..18, both items:')
print(f'  aready seen -> first MMA issued (weights)      {mean(lambda r: r[3] - r[2]):6.2f} us')
print(f'  first MMA issued -> all issued                 {mean(lambda r: r[4] - r[3]):6.2f} us')
if len(body[0]) > 10:
  print(f'    of which waiting for weight stages           {mean(lambda r: r[10]):6.2f} us')
print(f'  all issued -> epilogue sees tfull              {mean(lambda r: r[6] - r[4]):6.2f} us')
print(f'  epilogue waits for tfull                       {mean(lambda r: r[6] - r[5]):6.2f} us')
print(f'  tfull -> accumulator + residual in registers   {mean(lambda r: r[7] - r[6]):6.2f} us')
print(f'  registers -> operand written                   {mean(lambda r: r[8] - r[7]):6.2f} us')
print(f'  operand written -> arrived                     {mean(lambda r: r[9] - r[8]):6.2f} us')
nxt = {(r[0], r[1]): r for r in rows}
d = [us(nxt[(r[0] + 1, r[1])][2] - r[9]) for r in body if (r[0] + 1, r[1]) in nxt]
print(f'  arrived -> MMA thread sees aready (next round) {st.mean(d):6.2f} us')
d = [us(nxt[(r[0] + 1, r[1])][2] - r[2]) for r in body if (r[0] + 1, r[1]) in nxt]
print(f'  period of an item (aready to aready)           {st.mean(d):6.2f} us')
