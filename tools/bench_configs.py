"""Throughput of the other BASELINE.json configs (bench.py measures configs[1] = c2 only):

  c1  RNA 5'UTR MRL SVDD-MC, L=50, B=10, M=10           (the reference's CPU-runnable case)
  c3  DNA enhancer SVDD-PM (Tweedie x0 + Enformer-style oracle), L=200, B=128, M=10
  c4  DNA SVDD-MC, M=20: the per-GPU shard of batch 4096 on 8 GPUs (B=512); --full adds B=4096
  c5  RNA SVDD-PM, L=50, M=50, alpha in {0, 0.1, 1}: the per-GPU shard of batch 8192 (B=1024)

    python tools/bench_configs.py [--configs c1,c3,c4,c5] [--reps 2] [--out gpurun_out/configs.jsonl]

One JSON line per case: decoded seq/s (device-timed CUDA events over `reps` replays of the
captured 128-step trajectory after one warm-up run that includes the capture), ms per reverse
step, nominal TFLOP per step and the fraction of the measured sustained bf16 peak.
Random-init weights of the named architectures, in-kernel Philox noise."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench    # noqa: E402
from svdd_b200 import synthetic as helpers  # noqa: E402

F_VAL_ENF = 3.362e9
F_VAL_GRU = 17.18e6


def f_den(L):
  return 2 * L * (5 * 128 * 9 + 20 * 128 * 128 * 9 + 128 * 128 + 128 * 5)


def rna_models(device, oracle):
  from svdd_b200 import config, diffusion_gosai
  cfg = config.load_config('rna')
  torch.manual_seed(44)
  model = diffusion_gosai.Diffusion(cfg).to(device).eval()
  emb, head = helpers.build_convgru_oracle() if oracle else helpers.build_convgru_value()
  return model, emb.to(device), head.to(device)


def timed(run, reps):
  run()                                   # eager sizing pass + graph capture + first replay
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    x = run()
  e1.record()
  torch.cuda.synchronize()
  assert int(x.max()) <= 3 and int(x.min()) >= 0
  return e0.elapsed_time(e1) / reps


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--configs', default='c1,c3,c4,c5')
  ap.add_argument('--reps', type=int, default=2)
  ap.add_argument('--full', action='store_true', help='also run c4 with the whole batch of 4096 on this GPU')
  ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'configs.jsonl'))
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  from svdd_b200 import value_nets
  pk = bench.peaks()
  want = set(args.configs.split(','))
  cases = []
  if 'c1' in want:
    m, e, h = rna_models(dev, oracle=False)
    cases.append(('c1 RNA SVDD-MC L=50 B=10 M=10', 10, 10, 50, 10 * f_den(50) + 100 * F_VAL_GRU,
                  lambda m=m, e=e, h=h: m.controlled_sample(e, h, eval_sp_size=10, sample_M=10)))
  if want & {'c3', 'c4'}:
    _, dm, de, dh = bench.build_models(dev)
    if 'c3' in want:
      rm = value_nets.OriBaseModel(de, dh)
      cases.append(('c3 DNA SVDD-PM L=200 B=128 M=10', 128, 10, 200,
                    128 * 11 * f_den(200) + 1280 * F_VAL_ENF,
                    lambda: dm.controlled_sample_tweedie(rm, eval_sp_size=128, sample_M=10,
                                                         options='True', task='dna')))
    if 'c4' in want:
      for B in (512, 4096) if args.full else (512,):
        cases.append((f'c4 DNA SVDD-MC L=200 B={B} M=20' + (' (shard of 8)' if B == 512 else ' (whole batch)'),
                      B, 20, 200, B * f_den(200) + B * 20 * F_VAL_ENF,
                      lambda B=B: dm.controlled_sample(de, dh, eval_sp_size=B, sample_M=20)))
  if 'c5' in want:
    m5, e5, h5 = rna_models(dev, oracle=True)
    rm5 = value_nets.OriBaseModel(e5, h5)
    for alpha in (0.0, 0.1, 1.0):
      cases.append((f'c5 RNA SVDD-PM L=50 B=1024 M=50 alpha={alpha} (shard of 8)', 1024, 50, 50,
                    1024 * 51 * f_den(50) + 1024 * 50 * F_VAL_GRU,
                    lambda a=alpha: m5.controlled_sample_tweedie(rm5, eval_sp_size=1024, sample_M=50,
                                                                 options='True', task='rna', alpha=a)))
  os.makedirs(os.path.dirname(args.out), exist_ok=True)
  with open(args.out, 'a') as f:
    for name, B, M, L, flops_step, run in cases:
      ms = timed(run, args.reps)
      step_ms = ms / 128
      tf = flops_step / (step_ms * 1e-3) / 1e12
      line = {'workload': name, 'B': B, 'M': M, 'L': L, 'denoise_steps': 128,
              'decoded_seqs_per_sec': B / (ms * 1e-3), 'ms_per_decode': ms,
              'ms_per_denoise_step': step_ms, 'nominal_tflop_per_step': flops_step / 1e12,
              'achieved_tflops': tf, 'frac_of_sustained_bf16_peak': tf / pk['tf_sust'],
              'peak_source': pk['src'], 'reps': args.reps, 'n_gpus': 1,
              'noise': 'in-kernel Philox4x32-10', 'data': 'synthetic'}
      print(json.dumps(line), flush=True)
      f.write(json.dumps(line) + '\n')
      torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
