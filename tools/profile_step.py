"""Runs a few EAGER reverse steps of the bench workload (no CUDA graph) so that ncu can
list / profile individual launches:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 1
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--warm', type=int, default=1)
ap.add_argument('--B', type=int, default=bench.B_PER_GPU)
ap.add_argument('--M', type=int, default=bench.M)
ap.add_argument('--mode', default='mc')
args = ap.parse_args()
dev = torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
model.use_cuda_graph = False
from svdd_b200 import value_nets
if args.mode == 'mc':
  run = lambda n: model.controlled_sample(emb, head, num_steps=n, eval_sp_size=args.B, sample_M=args.M)
else:
  rm = value_nets.OriBaseModel(emb, head)
  run = lambda n: model.controlled_sample_tweedie(rm, num_steps=n, eval_sp_size=args.B, sample_M=args.M,
                                                  options='True', task='dna')
run(args.warm)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('measured')
run(args.steps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print('done')
