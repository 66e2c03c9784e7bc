"""``BaseModel`` -- the orchestrator that ``decode.py`` / ``decode_tweedie.py`` drive
(reference: Enformer.py:21-864, decode path only).

Same constructor arguments and the same two entry points
  controlled_decode(gen_batch_num, sample_M)                 Enformer.py:400-477
  controlled_decode_tweedie(gen_batch_num, sample_M, options) Enformer.py:720-813
returning the reference's 5-tuple
  (samples, value_func_preds, reward_model_preds, top_k / selected baseline, baseline_preds).

Differences that are forced by the environment or are additions:
  * checkpoints: the reference hard-codes ``artifacts/{RNA,DNA}_Diffusion:v0`` and loads the
    reward oracles through gReLU; both are honoured when the files exist, and
    ``random_init=True`` builds the same architectures with seeded random weights (the
    W&B artifacts are not available offline);
  * scoring consumes token ids (the kernels fuse the one-hot), so ``transform_samples``
    is kept only as a utility;
  * when ``torch.distributed`` is initialised the batch is sharded over ranks
    (svdd_b200/sharding.py) and gathered at the end -- candidates never cross ranks.
Training (``forward``, optimizers) is out of scope.
"""
import os
import time

import torch
from torch import nn

from . import config as config_mod
from . import diffusion_gosai, sharding, value_nets

RNA_TASKS = ('rna', 'rna_saluki')


def _convgru_value_nets():
  emb = value_nets.ConvGRUTrunk(
      stem_in_channels=4, stem_channels=64, stem_kernel_size=15, n_conv=6, channel_init=64,
      channel_mult=1, kernel_size=5, act_func='relu', conv_norm=True, pool_func=None,
      pool_size=None, residual=True, crop_len=0, n_gru=1, dropout=0.1, gru_norm=True)
  head = value_nets.ConvHead(n_tasks=1, in_channels=64, act_func=None, pool_func='avg', norm=False)
  return emb, head                                                    # Enformer.py:32-49


def random_reward_model(task):
  """Random-init reward oracle of the architecture the reference loads from gReLU
  checkpoints: RNA = ConvGRUModel(n_conv=6, stem_channels=64, channel_init=64)
  (rna_MRL_oracle.py:38-44), DNA = Enformer-family regressor with 3 tasks, HepG2 first
  (oracle.py:72,228)."""
  if task in RNA_TASKS:
    emb = value_nets.ConvGRUTrunk(stem_in_channels=4, stem_channels=64, n_conv=6, channel_init=64)
    head = value_nets.ConvHead(n_tasks=1, in_channels=64, act_func=None, pool_func='avg')
  else:
    emb = value_nets.EnformerTrunk(n_conv=7, channels=1536, n_transformers=11, n_heads=8, key_len=64)
    head = value_nets.ConvHead(n_tasks=3, in_channels=2 * 1536, act_func=None, pool_func='avg')
    with torch.no_grad():      # zero-initialised to_out would switch the attention path off
      for blk in emb.transformer_tower.blocks:
        nn.init.normal_(blk.mha.to_out.weight, std=blk.mha.to_out.weight.shape[1] ** -0.5)
  return value_nets.OriBaseModel(emb, head)


def load_grelu_reward_model(path, task):
  """Best-effort reader of a gReLU ``LightningModel`` checkpoint (Enformer.py:104-131):
  ``state_dict`` keys ``model.embedding.*`` / ``model.head.*`` are mapped onto the containers
  of this package.  (Untested against the real artifacts: they are unavailable offline.)"""
  ckpt = torch.load(path, map_location='cpu', weights_only=False)
  sd = ckpt.get('state_dict', ckpt)
  emb_sd = {k.split('embedding.', 1)[1]: v for k, v in sd.items() if 'embedding.' in k}
  head_sd = {k.split('head.', 1)[1]: v for k, v in sd.items() if 'head.' in k and 'embedding.' not in k}
  model = random_reward_model(task)
  model.embedding.load_state_dict(emb_sd, strict=True)
  model.head.load_state_dict(head_sd, strict=True)
  return model


class BaseModel(nn.Module):
  def __init__(self, embedding, head, cdq, batch_size, val_batch_num, timed=False,
               task='rna_saluki', n_tasks=1, saluki_body=0, *, random_init=False,
               artifacts_dir='artifacts', config_dir=None, alpha=0.0, build_eval_batches=False):
    super().__init__()
    self.task, self.n_tasks, self.saluki_body = task, n_tasks, saluki_body
    if task == 'rna_saluki':
      raise NotImplementedError('rna_saluki needs a private .npy the reference does not ship')
    if timed:
      raise NotImplementedError('TimedEnformerTrunk is not selected by decode.py defaults')
    if task in RNA_TASKS:                               # Enformer.py:31-49
      self.embedding, self.head = _convgru_value_nets()
    else:
      self.embedding, self.head = embedding, head
    self.cdq, self.timed, self.alpha = cdq, timed, alpha
    self.NUM_SAMPLES_PER_BATCH = batch_size
    cfg = config_mod.load_config(task, config_dir)
    rna = task in RNA_TASKS
    ckpt = os.path.join(artifacts_dir, 'RNA_Diffusion:v0/best.ckpt' if rna else 'DNA_Diffusion:v0/last.ckpt')
    rckpt = os.path.join(artifacts_dir, 'RNA_evaluation:v0/model.ckpt' if rna else 'DNA_evaluation:v0/model.ckpt')
    if random_init:
      self.ref_model = diffusion_gosai.Diffusion(cfg)
      self.reward_model = random_reward_model(task)
    else:
      for p in (ckpt, rckpt):
        if not os.path.isfile(p):
          raise FileNotFoundError(
              f'{p} not found: download the W&B artifacts as in the reference README, or pass '
              'random_init=True / --random_init for synthetic weights')
      print('CKPT_PATH: ', ckpt)
      self.ref_model = diffusion_gosai.Diffusion.load_from_checkpoint(ckpt, config=cfg, map_location='cpu')
      self.reward_model = load_grelu_reward_model(rckpt, task)
    self.ref_model.eval()
    self.reward_model.eval()
    for p in list(self.ref_model.parameters()) + list(self.reward_model.parameters()):
      p.requires_grad = False
    self.val_data_num = val_batch_num * batch_size
    self._val_batch_num = val_batch_num
    self._build_eval_batches = build_eval_batches

  # -- checkpoints ---------------------------------------------------------------------------
  # Key families of a reference value-function checkpoint (``trainer.py:73-88`` saves
  # ``BaseModel.state_dict()`` under 'model_state_dict'; decode.py:101-104 loads it strict=True):
  #   embedding.* / head.*            the value net                        -> same names here
  #   ref_model.backbone.*            the frozen MDLM denoiser             -> same names here
  #   ref_model.<anything else>       Lightning metric / EMA state         -> ignored
  #   reward_model.model.embedding.* / reward_model.model.head.*
  #                                   gReLU LightningModel wraps its net in ``.model``
  #                                   (Enformer.py:104-131)                -> reward_model.embedding.* / .head.*
  #   reward_model.<anything else>    gReLU metrics / transforms           -> ignored
  @staticmethod
  def map_reference_keys(state_dict):
    """Reference-layout BaseModel state_dict -> (this package's layout, ignored keys)."""
    mapped, ignored = {}, []
    for k, v in state_dict.items():
      if k.startswith('reward_model.model.'):
        mapped['reward_model.' + k[len('reward_model.model.'):]] = v
      elif k.startswith('reward_model.') and not (k.startswith('reward_model.embedding.') or
                                                  k.startswith('reward_model.head.')):
        ignored.append(k)
      elif k.startswith('ref_model.') and not k.startswith('ref_model.backbone.'):
        ignored.append(k)
      else:
        mapped[k] = v
    return mapped, ignored

  def load_state_dict(self, state_dict, strict=True, assign=False):
    """Accepts the reference's own ``model_state_dict`` (key mapping above) as well as this
    package's.  strict=True keeps the reference's contract for everything that is on the
    decode path: every parameter of this model must be present and no unknown key may remain;
    only the documented bookkeeping families are dropped."""
    mapped, ignored = self.map_reference_keys(state_dict)
    result = super().load_state_dict(mapped, strict=strict, assign=assign)
    self.ignored_checkpoint_keys = ignored
    return result

  # -- utilities -----------------------------------------------------------------------------
  def transform_samples(self, samples, num_classes=4):
    return self.ref_model.transform_samples(samples, num_classes)

  def _local_batch(self):
    return sharding.local_rows(self.NUM_SAMPLES_PER_BATCH)

  def _reward(self, tokens):
    """reward_model(onehot.float().transpose(1, 2))[:, 0] on token ids (Enformer.py:447)."""
    return value_nets.score_tokens(self.reward_model.embedding, self.reward_model.head, tokens)

  def _value(self, tokens):
    """head(embedding(onehot.float())).squeeze(2) on token ids (Enformer.py:443)."""
    return value_nets.score_tokens(self.embedding, self.head, tokens)

  def build_eval_batches(self):
    """The reference's constructor rolls out val_batch_num plain samples with all
    intermediate states and their final rewards (Enformer.py:135-160; feeds value-function
    training / eval.py).  Not needed for decoding, so it is opt-in here."""
    steps = self.ref_model.config.sampling.steps
    xs = [[] for _ in range(steps)]
    ys = [[] for _ in range(steps)]
    for _ in range(self._val_batch_num):
      samples, mids = self.ref_model._sample(eval_sp_size=self.NUM_SAMPLES_PER_BATCH)
      target = self._reward(samples)
      for j, s in enumerate(mids + [samples]):
        xs[j].append(s)
        ys[j].append(target)
    self.eval_time_step_batches = [torch.cat(x, 0) for x in xs]
    self.eval_time_step_targets = [torch.cat(y, 0) for y in ys]

  # -- decoding ----------------------------------------------------------------------------------
  @staticmethod
  def _tick():
    if torch.cuda.is_available():
      torch.cuda.synchronize()
    return time.perf_counter()

  def _decode(self, gen_batch_num, sample_M, sampler, topk):
    """The body shared by controlled_decode / controlled_decode_tweedie (Enformer.py:439-477,
    762-813).  ``self.timing`` afterwards holds the wall-clock seconds of its two phases (one
    device synchronisation at each phase boundary): the value-weighted SVDD batches with their
    final value / reward scoring, and the gen_batch_num * sample_M plain baseline rollouts."""
    rows, offset = self._local_batch()
    samples, value_preds, reward_preds = [], [], []
    t0 = self._tick()
    for i in range(gen_batch_num):
      batch = sampler(rows, offset + i * self.NUM_SAMPLES_PER_BATCH)
      batch = sharding.gather_rows(batch)                # final gather of the sequences
      samples.append(batch)
      value_preds.append(self._value(batch))
      reward_preds.append(self._reward(batch))
    t1 = self._tick()
    print('Value-weighted sampling done.')
    baseline_preds, all_preds = [], []
    for i in range(gen_batch_num * sample_M):            # Enformer.py:456-467
      batch = self.ref_model.decode_sample(eval_sp_size=rows, row_offset=offset + (gen_batch_num + i) * self.NUM_SAMPLES_PER_BATCH)
      pred = self._reward(sharding.gather_rows(batch))
      if i < gen_batch_num:
        baseline_preds.append(pred)
      all_preds.append(pred)
    t2 = self._tick()
    print('Baseline sampling done.')
    self.timing = {'svdd_s': t1 - t0, 'baseline_s': t2 - t1, 'svdd_batches': gen_batch_num,
                   'baseline_rollouts': gen_batch_num * sample_M, 'batch_size': self.NUM_SAMPLES_PER_BATCH}
    baseline = torch.cat(baseline_preds)
    if topk:                                              # Enformer.py:471-475
      all_values = torch.cat(all_preds)
      top_k_values, _ = torch.topk(all_values, int(len(all_values) / sample_M))
    else:                                                 # Enformer.py:802
      top_k_values = baseline
    return samples, torch.cat(value_preds), torch.cat(reward_preds), top_k_values, baseline

  @torch.no_grad()
  def controlled_decode(self, gen_batch_num, sample_M):
    def sampler(rows, row_offset):
      return self.ref_model.controlled_sample(self.embedding, self.head, eval_sp_size=rows,
                                              sample_M=sample_M, alpha=self.alpha, row_offset=row_offset)
    return self._decode(gen_batch_num, sample_M, sampler, topk=True)

  @torch.no_grad()
  def controlled_decode_tweedie(self, gen_batch_num, sample_M, options):
    def sampler(rows, row_offset):
      return self.ref_model.controlled_sample_tweedie(
          self.reward_model, eval_sp_size=rows, sample_M=sample_M, options=options, task=self.task,
          alpha=self.alpha, row_offset=row_offset)
    samples, v, r, top, base = self._decode(gen_batch_num, sample_M, sampler, topk=False)
    return [row for batch in samples for row in batch], v, r, top, base    # .extend semantics, :763
