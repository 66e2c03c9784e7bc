import os, sys, torch
ROOT='/root/repo'
sys.path[:0]=[ROOT, ROOT+'/tests']
import helpers
from svdd_b200 import value_nets
dev=torch.device('cuda:0')
emb, head = helpers.build_convgru_oracle()
sc = value_nets.packed_scorer(emb.to(dev), head.to(dev))
x = torch.randint(0,4,(16384,50),dtype=torch.uint8,device=dev)
sc.score(x); torch.cuda.synchronize()
torch.cuda.nvtx.range_push('m'); sc.score(x); torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
