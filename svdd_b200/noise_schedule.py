"""Noise schedule of the decode path (reference: noise_schedule.py:126-152,
``LogLinearNoise``; selected by ``noise.type: loglinear`` in
configs_gosai/noise/loglinear.yaml).  Host-side fp32 scalar math only: the two
move chances per reverse step are computed here with the reference's expression
sequence and handed to the CUDA kernels as scalars (their low bits differ from
the closed form 0.999*t)."""
import torch
from torch import nn


class LogLinearNoise(nn.Module):
  """sigma(t) = -log1p(-(1 - eps) t), so that 1 - exp(-sigma(t)) = (1 - eps) t."""

  def __init__(self, eps=1e-3):
    super().__init__()
    self.eps = eps
    self.sigma_max = self.total_noise(torch.tensor(1.0))
    self.sigma_min = self.eps + self.total_noise(torch.tensor(0.0))

  def rate_noise(self, t):
    return (1 - self.eps) / (1 - (1 - self.eps) * t)

  def total_noise(self, t):
    return -torch.log1p(-(1 - self.eps) * t)

  def forward(self, t):
    return self.total_noise(t), self.rate_noise(t)


def get_noise(config, dtype=torch.float32):
  kind = config.noise.type
  if kind != 'loglinear':
    raise NotImplementedError(
        f"noise.type '{kind}': the decode path is built for 'loglinear' "
        '(configs_gosai/config_gosai.yaml defaults)')
  return LogLinearNoise()


def move_chance_schedule(noise, num_steps, eps, device='cpu'):
  """[(mc_t, mc_s, sigma_t, sigma_s)] per reverse step as python floats holding
  fp32 values, plus sigma at the last timestep for the noise-removal forward.

  Follows diffusion_gosai.py:1036-1043 (timesteps = linspace(1, eps, N+1);
  dt = (1 - eps)/N; t = timesteps[i] * ones(B, 1)) and :1176-1187
  (sigma = noise(t); mc = 1 - exp(-sigma)) on a [1, 1] tensor -- every sequence of
  a batch shares the values.  The product evaluates it on the CPU (what the oracle
  and the goldens pin); ``device`` exists so that tests can evaluate the same
  expression sequence with the CUDA libm and compare the low bits
  (tests/test_gpu_sampling.py::test_schedule_host_vs_device)."""
  timesteps = torch.linspace(1, eps, num_steps + 1, device=device)
  dt = (1 - eps) / num_steps
  rows = []
  for i in range(num_steps):
    t = timesteps[i] * torch.ones(1, 1, device=device)
    sigma_t = noise(t)[0].squeeze(-1)
    sigma_s = noise(t - dt)[0].squeeze(-1)
    mc_t = 1 - torch.exp(-sigma_t)
    mc_s = 1 - torch.exp(-sigma_s)
    rows.append((mc_t.item(), mc_s.item(), sigma_t.item(), sigma_s.item()))
  sigma_last = noise(timesteps[-1] * torch.ones(1, 1, device=device))[0].item()
  return rows, sigma_last
