"""Config loading for the decode path.  The reference composes a hydra tree
(``Enformer.py:78-89``: ``configs_gosai`` for DNA, ``configs_gosai_rna`` for RNA)
and the decode path reads a dozen keys of it; hydra/omegaconf are not required
here -- ``svdd_b200/configs/{dna,rna}.yaml`` carry exactly those keys and load
into attribute namespaces.  A reference config directory can be passed instead."""
import os
import types

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))


def _ns(obj):
  if isinstance(obj, dict):
    return types.SimpleNamespace(**{k: _ns(v) for k, v in obj.items()})
  return obj


# configs_gosai/model/small.yaml, the DiT the reference's `backbone: dit` branch would build
# (`length` is the SEQUENCE length the sampler uses, so the task's value is kept)
DIT_SMALL = dict(name='small', type='ddit', hidden_size=768, cond_dim=128, n_blocks=12, n_heads=12,
                 scale_by_sigma=True, dropout=0.1, tie_word_embeddings=False)


def load_config(task='dna', config_dir=None, backbone=None, **overrides):
  """task: 'dna' (anything that is not rna/rna_saluki, as in Enformer.py:75-89) or 'rna'.
  backbone='dit' swaps the model section for configs_gosai/model/small.yaml (hydra:
  ``model=small backbone=dit``)."""
  name = 'rna' if task in ('rna', 'rna_saluki') else 'dna'
  if config_dir is None:
    with open(os.path.join(_HERE, 'configs', name + '.yaml')) as f:
      tree = yaml.safe_load(f)
  else:
    # a reference-style directory: config_gosai.yaml + model/dnaconv.yaml
    with open(os.path.join(config_dir, 'config_gosai.yaml')) as f:
      tree = yaml.safe_load(f)
    with open(os.path.join(config_dir, 'model', 'dnaconv.yaml')) as f:
      tree['model'] = yaml.safe_load(f)
    tree.setdefault('noise', {'type': 'loglinear'})
    if not isinstance(tree['noise'], dict):
      tree['noise'] = {'type': 'loglinear'}
    tree['loader'] = {'eval_batch_size': 512}
  if backbone == 'dit':
    tree['backbone'] = 'dit'
    tree['model'] = dict(DIT_SMALL, length=tree['model']['length'])
  elif backbone is not None:
    tree['backbone'] = backbone
  for dotted, v in overrides.items():
    node = tree
    keys = dotted.split('.')
    for k in keys[:-1]:
      node = node[k]
    node[keys[-1]] = v
  return _ns(tree)
