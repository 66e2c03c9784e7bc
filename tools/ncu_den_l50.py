"""One fused-denoiser launch at L=50 (two sequences per CTA, split epilogue) for an ncu capture."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers  # noqa: E402

dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5920
den = helpers.build_denoiser(44, 50).to(dev).packed()
x = helpers.random_tokens(n, 50, 3, 0.5).to(dev).to(torch.uint8)
out = torch.empty((n, 50, 5), device=dev)
den.forward(x, 0.0, out=out)
torch.cuda.synchronize()
den.forward(x, 0.0, out=out)
torch.cuda.synchronize()
