"""SVDD-MC decoding CLI -- drop-in for the reference's ``decode.py`` (flags at
decode.py:124-211, run() at :52-121): same ``--task/--sample_M/--batch_size/--val_batch_num/
--seed/--reward_name/--load_checkpoint_path/--model`` flags, same output
``./log/{task}-{reward_name}.npz`` with float32 arrays ``decoding`` and ``baseline``.

Added flags: ``--random_init`` (synthetic weights; the W&B artifacts are unavailable
offline), ``--alpha`` (soft selection; 0 = the reference's argmax), ``--artifacts_dir``,
``--out_dir``.  Multi-GPU: launch with torchrun; the batch is sharded over ranks."""
import argparse
import os
import random

import numpy as np
import torch


def set_seed(seed):
  random.seed(seed)
  np.random.seed(seed)
  torch.manual_seed(seed)
  torch.cuda.manual_seed_all(seed)


def build_parser(tweedie=False):
  p = argparse.ArgumentParser()
  p.add_argument('--run_name', type=str, default='decode')
  p.add_argument('--debug', action='store_true')
  p.add_argument('--task', type=str, default='rna_saluki' if tweedie else 'DNA')
  p.add_argument('--saluki_body', type=int, default=0)
  p.add_argument('--n_task', type=int, default=1)
  p.add_argument('--model', type=str, default='enformer')
  p.add_argument('--batch_size', type=int, default=256)
  p.add_argument('--sample_M', type=int, default=20 if tweedie else 5)
  p.add_argument('--val_batch_num', type=int, default=1)
  p.add_argument('--seed', type=int, default=44)
  p.add_argument('--reward_name', type=str, default='HepG2')
  p.add_argument('--load_checkpoint_path', type=str, default=None)
  p.add_argument('--pre_model_path', default=None)
  p.add_argument('--cdq', action='store_true')
  p.add_argument('--dist', action='store_true')
  if tweedie:
    p.add_argument('--tweedie', type=str, default=True, required=True,
                   help='Use Tweedie formula ("True") or the masked-one-hot heuristic')
  # accepted for command-line compatibility with the reference, unused on the decode path
  for name, typ, default in (('scaffold', None, None), ('lstm', None, None), ('data_name', str, 'moses2'),
                             ('num_props', int, 0), ('tokenizer', str, 'simple'), ('n_layer', int, 8),
                             ('n_head', int, 8), ('n_embd', int, 768), ('max_epochs', int, 1),
                             ('max_iters', int, 50000), ('num_workers', int, 12),
                             ('save_start_epoch', int, 120), ('save_interval_epoch', int, 10),
                             ('learning_rate', float, 6e-4), ('lstm_layers', int, 0), ('max_len', int, 512),
                             ('grad_norm_clip', float, 1.0), ('auto_fp16to32', None, None),
                             ('pre_root_path', str, None), ('root_path', str, None),
                             ('output_tokenizer_dir', str, None), ('fix_condition', str, None),
                             ('conditions_path', str, None), ('conditions_split_id_path', str, None)):
    if typ is None:
      p.add_argument('--' + name, action='store_true')
    else:
      p.add_argument('--' + name, type=typ, default=default)
  p.add_argument('--props', nargs='+', default=['qed'])
  # additions
  p.add_argument('--random_init', action='store_true', help='seeded random weights instead of checkpoints')
  p.add_argument('--alpha', type=float, default=0.0, help='soft selection temperature (0 = argmax, reference)')
  p.add_argument('--artifacts_dir', type=str, default='artifacts')
  p.add_argument('--out_dir', type=str, default='./log')
  return p


def run(args, tweedie=False):
  from svdd_b200 import sharding
  from svdd_b200.base_model import BaseModel
  from svdd_b200.value_nets import ConvHead, EnformerTrunk
  if 'LOCAL_RANK' in os.environ and int(os.environ.get('WORLD_SIZE', 1)) > 1:
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    torch.distributed.init_process_group('nccl')
  set_seed(args.seed)
  print('loading model')
  if args.model != 'enformer':
    raise NotImplementedError(f"--model {args.model}: decode.py's default 'enformer' is the one built")
  trunk = EnformerTrunk(n_conv=7, channels=1536, n_transformers=11, n_heads=8, key_len=64,
                        attn_dropout=0.05, pos_dropout=0.01, ff_dropout=0.4, crop_len=0)
  head = ConvHead(n_tasks=1, in_channels=2 * 1536, act_func=None, pool_func='avg')
  model = BaseModel(embedding=trunk, head=head, cdq=args.cdq, batch_size=args.batch_size,
                    val_batch_num=1, task=args.task, n_tasks=args.n_task, saluki_body=args.saluki_body,
                    random_init=args.random_init, artifacts_dir=args.artifacts_dir, alpha=args.alpha)
  for path in (args.pre_model_path, args.load_checkpoint_path):
    if path is not None:
      print('loading stored model: ', path)
      ckpt = torch.load(path, map_location='cpu', weights_only=False)
      model.load_state_dict(ckpt['model_state_dict'], strict=True)
  print('total params:', sum(p.numel() for p in model.parameters()))
  model.cuda()
  model.eval()
  if tweedie:
    out = model.controlled_decode_tweedie(gen_batch_num=args.val_batch_num, sample_M=args.sample_M,
                                          options=args.tweedie)
  else:
    out = model.controlled_decode(gen_batch_num=args.val_batch_num, sample_M=args.sample_M)
  _, value_func_preds, reward_model_preds, selected_baseline_preds, baseline_preds = out
  ours = reward_model_preds.float().cpu().numpy()
  baseline = baseline_preds.float().cpu().numpy()
  if sharding.world()[0] == 0:
    os.makedirs(args.out_dir, exist_ok=True)
    name = '%s-%s' % (args.task, args.reward_name) + ('_tw' if tweedie else '')
    np.savez(os.path.join(args.out_dir, name), decoding=ours, baseline=baseline)
    print('decoding median %.4f  baseline median %.4f  -> %s.npz' %
          (float(np.median(ours)), float(np.median(baseline)), os.path.join(args.out_dir, name)))
    t = model.timing
    n = t['batch_size']
    print('timing: SVDD %d x %d sequences in %.3f s (%.1f seq/s incl. capture and final scoring); '
          'baseline %d x %d rollouts in %.3f s (%.1f seq/s)' %
          (t['svdd_batches'], n, t['svdd_s'], t['svdd_batches'] * n / t['svdd_s'],
           t['baseline_rollouts'], n, t['baseline_s'], t['baseline_rollouts'] * n / t['baseline_s']))
  return out


if __name__ == '__main__':
  run(build_parser().parse_args())
