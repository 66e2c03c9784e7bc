"""Batch sharding across the GPUs of one box.

Every sequence's trajectory depends only on its own state, its own M candidates and its
own noise (diffusion_gosai.py:1203-1227 is row-wise), so the path shards over the batch
with NO per-step communication: rank r owns the contiguous rows
[r*ceil(B/G), min(B, (r+1)*ceil(B/G))).  The in-kernel noise is keyed by the GLOBAL row
index (``row_offset``), so results do not depend on the number of ranks.  The only
collective is the final gather of the decoded rows (NCCL ``all_gather`` over NVLink; gloo
in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def partition(B, rank, world_size):
  """Contiguous block partition of B rows -> (row_offset, n_rows) of `rank`."""
  per = -(-B // world_size)
  lo = min(B, rank * per)
  hi = min(B, lo + per)
  return lo, hi - lo


def local_rows(B):
  """(n_rows, row_offset) of this rank for a global batch of B rows."""
  rank, ws = world()
  lo, n = partition(B, rank, ws)
  return n, lo


def gather_rows(x, B=None):
  """All ranks contribute their [n_r, ...] block; every rank receives the [B, ...]
  concatenation in row order.  Ragged last blocks are padded for the collective."""
  rank, ws = world()
  if ws == 1:
    return x
  n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
  counts = [torch.zeros_like(n) for _ in range(ws)]
  dist.all_gather(counts, n)
  counts = [int(c) for c in counts]
  mx = max(counts)
  if x.shape[0] < mx:
    pad = torch.zeros((mx - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    x = torch.cat([x, pad], 0)
  parts = [torch.empty_like(x) for _ in range(ws)]
  dist.all_gather(parts, x.contiguous())
  return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
