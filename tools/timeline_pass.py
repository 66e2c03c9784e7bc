"""CTA-0 pipeline milestones of every gemm2 launch of one warm c2 value pass (SVDD_TIMELINE=1 synchronises after
every launch and prints clock64 deltas: setup, first TMA, first tile's MMAs done, stores done, exit):
    SVDD_TIMELINE=1 python tools/timeline_pass.py
Round-2 reading: ~1.7 us of set-up, first TMA at ~2.7 us, the first (cold) tile of a 1x1 launch takes ~7 us instead of
4.4, ~0.9 us of tear-down: 8-9 us of fixed cost per launch, ~0.2 ms per c2 step over the 26 launches of the conv tower."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT]
import bench
from svdd_b200 import value_nets
dev=torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
cand=torch.randint(0,4,(1280,200),device=dev,dtype=torch.uint8)
sc=value_nets.packed_scorer(emb, head)
out=torch.empty(1280,device=dev)
sc.score(cand,out=out); torch.cuda.synchronize()
print('---- second pass', file=sys.stderr)
sc.score(cand,out=out); torch.cuda.synchronize()
