"""A few EAGER reverse steps of an RNA SVDD-PM run shaped like BASELINE config 5 (L = 50, M = 50,
Tweedie x0 + ConvGRU oracle) at B = 256 per launch, for ncu launch lists / captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "measured/" \
        --csv --log-file gpurun_out/launches_c5.csv python tools/profile_step_rna.py --steps 1"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from svdd_b200 import config, diffusion_gosai, synthetic, value_nets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--B', type=int, default=256)
ap.add_argument('--M', type=int, default=50)
ap.add_argument('--mc', action='store_true', help='SVDD-MC with the ConvGRU value net (BASELINE config 1: --mc --B 10 --M 10)')
args = ap.parse_args()
dev = torch.device('cuda:0')
torch.manual_seed(44)
model = diffusion_gosai.Diffusion(config.load_config('rna')).to(dev).eval()
model.use_cuda_graph = False
oe, oh = synthetic.build_convgru_oracle()
rm = value_nets.OriBaseModel(oe.to(dev), oh.to(dev))
if args.mc:
  ve, vh = synthetic.build_convgru_value()
  ve, vh = ve.to(dev), vh.to(dev)
  run = lambda n: model.controlled_sample(ve, vh, num_steps=n, eval_sp_size=args.B, sample_M=args.M)
else:
  run = lambda n: model.controlled_sample_tweedie(rm, num_steps=n, eval_sp_size=args.B, sample_M=args.M,
                                                  options='True', task='rna')
run(1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('measured')
run(args.steps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print('done')
