"""
Multi-GPU composition of the scoring path: the posterior state is tiny (<= 200 MB at
N = 5000) and is rebuilt redundantly on every rank; candidates -- independent given the posterior --
are sharded by contiguous global row ranges, one process per GPU; the ONLY collective is an
all-gather of one 16-byte (score, global index) pair per rank, reduced locally in np.argmax order
(first index on ties, NaN counts as the maximum: oper_utils.py:73) so that the result is identical
to a single np.argmax over the concatenated candidate set, whatever the number of ranks.

torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(m, rank, world):
  """ Contiguous balanced split of rows [0, m): rank r owns [m*r // world, m*(r+1) // world). """
  return (m * rank) // world, (m * (rank + 1)) // world


def better(score_a, index_a, score_b, index_b):
  """ True if (score_a, index_a) precedes (score_b, index_b) in np.argmax order.  A negative index
      marks 'no candidate'. """
  if index_b < 0:
    return index_a >= 0
  if index_a < 0:
    return False
  nan_a, nan_b = np.isnan(score_a), np.isnan(score_b)
  if nan_a or nan_b:
    if nan_a and nan_b:
      return index_a < index_b
    return bool(nan_a)
  if score_a > score_b:
    return True
  if score_a < score_b:
    return False
  return index_a < index_b


def reduce_pairs(scores, indices):
  """ Lexicographic (score desc, index asc, NaN first) reduction of per-rank winners. """
  best_s, best_i = 0.0, -1
  for s, i in zip(scores, indices):
    if better(float(s), int(i), best_s, best_i):
      best_s, best_i = float(s), int(i)
  return best_s, best_i


def all_reduce_argmax(score, global_index, device=None, group=None):
  """ One all-gather of a packed (fp64 score bits, int64 index) pair per rank + a local reduce. """
  if not (dist.is_available() and dist.is_initialized()):
    return float(score), int(global_index)
  world = dist.get_world_size(group)
  dev = torch.device('cpu') if device is None else device
  mine = torch.empty(2, dtype=torch.int64, device=dev)
  mine[0] = int(np.float64(score).view(np.int64))
  mine[1] = int(global_index)
  gathered = [torch.empty(2, dtype=torch.int64, device=dev) for _ in range(world)]
  dist.all_gather(gathered, mine, group=group)
  stacked = torch.stack(gathered).cpu().numpy()
  scores = stacked[:, 0].copy().view(np.float64)
  return reduce_pairs(scores, stacked[:, 1])


def all_reduce_argmax_point(score, global_index, point, dim, device=None, group=None):
  """ all_reduce_argmax with the winner's coordinates riding along: still ONE all-gather, of (2 + dim) 8-byte
      words per rank, so that a rank need not hold (or re-draw) candidate rows outside its own shard.  `point` may
      be None on a rank without candidates (index < 0).  Returns (score, index, point (dim,) ndarray). """
  if not (dist.is_available() and dist.is_initialized()):
    return float(score), int(global_index), point
  world = dist.get_world_size(group)
  dev = torch.device('cpu') if device is None else device
  words = np.zeros(2 + int(dim), dtype=np.int64)
  words[0] = int(np.float64(score).view(np.int64))
  words[1] = int(global_index)
  if point is not None:
    words[2:] = np.ascontiguousarray(np.asarray(point, dtype=np.float64)).view(np.int64)
  mine = torch.from_numpy(words).to(dev)
  gathered = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(gathered, mine, group=group)
  stacked = torch.stack(gathered).cpu().numpy()
  scores = stacked[:, 0].copy().view(np.float64)
  best_s, best_i = reduce_pairs(scores, stacked[:, 1])
  if best_i < 0:
    return best_s, best_i, None
  row = int(np.nonzero(stacked[:, 1] == best_i)[0][0])
  return best_s, best_i, stacked[row, 2:].copy().view(np.float64)


def all_reduce_argmax_many(scores, global_indices, device=None, group=None):
  """ K independent arg-maxes at once (Add-UCB groups, Thompson draws): one all-gather of K packed
      (score bits, index) pairs per rank, K local reductions.  Returns (scores (K,), indices (K,)). """
  scores = np.asarray(scores, dtype=np.float64).reshape(-1)
  idx = np.asarray(global_indices, dtype=np.int64).reshape(-1)
  if not (dist.is_available() and dist.is_initialized()):
    return scores.copy(), idx.copy()
  world = dist.get_world_size(group)
  dev = torch.device('cpu') if device is None else device
  mine = torch.from_numpy(np.stack((scores.view(np.int64), idx), axis=1).copy()).to(dev)
  gathered = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(gathered, mine, group=group)
  stacked = torch.stack(gathered).cpu().numpy()                 # (world, K, 2)
  out_s, out_i = np.empty(len(scores)), np.empty(len(scores), dtype=np.int64)
  for k in range(len(scores)):
    out_s[k], out_i[k] = reduce_pairs(stacked[:, k, 0].copy().view(np.float64), stacked[:, k, 1])
  return out_s, out_i


def sharded_score_argmax(score_fn, m_total, device=None, group=None):
  """ score_fn(lo, hi) -> (best_score, best_local_index) over global rows [lo, hi).  Returns the
      global (score, index) on every rank. """
  rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
  world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
  lo, hi = shard_bounds(m_total, rank, world)
  if hi > lo:
    s, i = score_fn(lo, hi)
    s, i = float(s), int(i) + lo
  else:
    s, i = 0.0, -1
  return all_reduce_argmax(s, i, device=device, group=group)
