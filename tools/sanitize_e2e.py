"""Small DNA SVDD-MC / SVDD-PM decodes (full-size networks, B = 2, M = 3, 3 steps, eager) for
`compute-sanitizer --tool memcheck python tools/sanitize_e2e.py`: every kernel of the hot path once."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
from svdd_b200 import value_nets  # noqa: E402

dev = torch.device('cuda:0')
cfg, model, emb, head = bench.build_models(dev)
model.use_cuda_graph = False
x = model.controlled_sample(emb, head, num_steps=3, eval_sp_size=2, sample_M=3)
print('mc', x.shape, int(x.max()))
rm = value_nets.OriBaseModel(emb, head)
x = model.controlled_sample_tweedie(rm, num_steps=3, eval_sp_size=2, sample_M=3, options='True', task='dna')
print('pm', x.shape, int(x.max()))
torch.cuda.synchronize()
# RNA: short-sequence denoiser (two items in flight from 297 sequences up), conv-stack + GRU kernels
from svdd_b200 import config, diffusion_gosai, synthetic  # noqa: E402
torch.manual_seed(44)
m = diffusion_gosai.Diffusion(config.load_config('rna')).to(dev).eval()
m.use_cuda_graph = False
emb, head = synthetic.build_convgru_value()
x = m.controlled_sample(emb.to(dev), head.to(dev), num_steps=3, eval_sp_size=int(os.environ.get('SAN_B', '40')), sample_M=10)
print('rna mc', x.shape, int(x.max()))
oe, oh = synthetic.build_convgru_oracle()
x = m.controlled_sample_tweedie(value_nets.OriBaseModel(oe.to(dev), oh.to(dev)), num_steps=2, eval_sp_size=700, sample_M=10,
                                options='True', task='rna')
print('rna pm', x.shape, int(x.max()))
torch.cuda.synchronize()
