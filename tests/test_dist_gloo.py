"""
The N > 1 path on CPU: two gloo processes shard a candidate set, each finds its local winner (with a
NumPy stand-in for the device scorer -- the collective logic is what is under test), and the
16-byte all-gather + lexicographic reduce must return exactly np.argmax of the whole set, including
ties across the shard boundary and NaNs.
"""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, out_dir):
  sys.path.insert(0, ROOT)
  import torch.distributed as dist
  from dragonfly_b200 import dist as D
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  scores = np.load(os.path.join(out_dir, 'scores_%s.npy' % case))

  def score_fn(lo, hi):
    local = scores[lo:hi]
    idx = int(np.argmax(local))
    return local[idx], idx
  s, i = D.sharded_score_argmax(score_fn, len(scores))
  np.save(os.path.join(out_dir, 'res_%s_%d.npy' % (case, rank)), np.array([s, float(i)]))
  dist.barrier()
  dist.destroy_process_group()


CASES = {
  'plain': lambda rs: rs.standard_normal(1001),
  'tie_across_shards': lambda rs: np.where(np.arange(1000) % 400 == 150, 7.0, rs.uniform(0, 1, 1000)),
  'nan_in_second_shard': lambda rs: np.where(np.arange(1000) == 777, np.nan, rs.uniform(0, 1, 1000)),
  'nan_in_both': lambda rs: np.where((np.arange(1000) == 777) | (np.arange(1000) == 30), np.nan,
                                     rs.uniform(0, 1, 1000)),
  'tiny': lambda rs: np.array([3.0]),
}


@pytest.mark.parametrize('case', sorted(CASES.keys()))
def test_two_rank_argmax_equals_numpy(tmp_path, case):
  scores = CASES[case](np.random.RandomState(0))
  np.save(os.path.join(str(tmp_path), 'scores_%s.npy' % case), scores)
  port = 29500 + (os.getpid() % 2000) + sorted(CASES.keys()).index(case)
  mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
  want = int(np.argmax(scores))
  for r in range(2):
    s, i = np.load(os.path.join(str(tmp_path), 'res_%s_%d.npy' % (case, r)))
    assert int(i) == want
    assert (np.isnan(s) and np.isnan(scores[want])) or s == scores[want]


def test_shard_bounds_cover_everything():
  from dragonfly_b200 import dist as D
  for m in [0, 1, 7, 1000, 4000001]:
    for world in [1, 2, 3, 8]:
      edges = [D.shard_bounds(m, r, world) for r in range(world)]
      assert edges[0][0] == 0 and edges[-1][1] == m
      assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
      sizes = [hi - lo for lo, hi in edges]
      assert max(sizes) - min(sizes) <= 1


def test_reduce_pairs_order():
  from dragonfly_b200 import dist as D
  assert D.reduce_pairs([1.0, 2.0, 2.0], [5, 9, 3]) == (2.0, 3)
  s, i = D.reduce_pairs([1.0, np.nan, np.nan], [0, 8, 4])
  assert np.isnan(s) and i == 4
  assert D.reduce_pairs([0.0, 5.0], [-1, 2]) == (5.0, 2)
  assert D.reduce_pairs([], []) == (0.0, -1)


def _gather_worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  import torch.distributed as dist
  from dragonfly_b200 import hp_grid
  from dragonfly_b200 import dist as D
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  total = 7
  lo, hi = D.shard_bounds(total, rank, world)
  mine = np.arange(lo, hi, dtype=np.float64) * -1.5 - 3.0        # stand-in LML values
  lmls, probs = hp_grid.gather_shards(mine, total)
  np.save(os.path.join(out_dir, 'gather_%d.npy' % rank), np.stack([lmls, probs]))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_hp_grid_gather(tmp_path):
  """ The hp-grid all-gather (uneven shards 3 + 4) and the rand_exp_sampling weights. """
  port = 31500 + (os.getpid() % 2000)
  mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  want = np.arange(7, dtype=np.float64) * -1.5 - 3.0
  w = np.exp(want - want.max()); w /= w.sum()
  for r in range(2):
    lmls, probs = np.load(os.path.join(str(tmp_path), 'gather_%d.npy' % r))
    assert (lmls == want).all()
    assert np.allclose(probs, w, rtol=1e-15)


# ---- the acquisition operators themselves under two ranks (host logic; NumPy stand-in for the device scorer) ----
def _acq_worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  from argparse import Namespace
  import torch.distributed as dist
  from dragonfly_b200 import gpb_acquisitions as A
  from dragonfly_b200.domains import EuclideanDomain
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  calls = []

  def scorer(pts):                                   # what gp._fused_score returns: (best, index, scores)
    vals = np.sin(7.0 * pts[:, 0]) + pts[:, 1] ** 2
    calls.append(len(pts))
    i = int(np.argmax(vals))
    return vals[i], i, None
  anc = Namespace(domain=EuclideanDomain([[0, 1], [-1, 2]]), max_evals=1001)
  np.random.seed(3)
  A.STREAM_SLAB_ROWS = 300          # four streamed slabs, the shard boundary (row 500) inside the second one
  pt = A._fused_maximise(scorer, anc)
  after = np.random.random()        # the global stream must be where one np.random.random((1001, 2)) leaves it
  # a sample vector with this rank's blocks filled in and the others at -inf (asy_ts)
  full = np.random.RandomState(5).standard_normal(10000)
  blk = 4096
  lo, hi = [(0, 8192), (8192, 10000)][rank] if world == 2 else (0, 10000)
  mine = np.full(10000, -np.inf); mine[lo:hi] = full[lo:hi]
  ts_idx = A._argmax_of_sharded_sample(mine)
  np.save(os.path.join(out_dir, 'acq_%d.npy' % rank), np.concatenate((pt, [sum(calls), ts_idx, after])))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_fused_maximise_equals_single_process(tmp_path):
  """ gpb_acquisitions._fused_maximise under 2 ranks: each rank scores half of the SAME seeded candidates,
      one 16-byte all-gather, identical recommendation on both ranks and to the single-process run. """
  from argparse import Namespace
  from dragonfly_b200 import gpb_acquisitions as A
  from dragonfly_b200.domains import EuclideanDomain
  port = 33500 + (os.getpid() % 2000)
  mp.spawn(_acq_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  np.random.seed(3)
  anc = Namespace(domain=EuclideanDomain([[0, 1], [-1, 2]]), max_evals=1001)

  def scorer(pts):
    vals = np.sin(7.0 * pts[:, 0]) + pts[:, 1] ** 2
    i = int(np.argmax(vals))
    return vals[i], i, None
  want = A._fused_maximise(scorer, anc)               # not distributed here: scores all 1001 rows, one slab
  want_after = np.random.random()
  np.random.seed(3)
  pts = A.draw_candidates(anc.domain.bounds, 1001)    # the reference's single draw (oper_utils.py:59-67)
  assert (want == pts[int(np.argmax(np.sin(7.0 * pts[:, 0]) + pts[:, 1] ** 2))]).all()
  want_ts = int(np.argmax(np.random.RandomState(5).standard_normal(10000)))
  sizes = []
  for r in range(2):
    got = np.load(os.path.join(str(tmp_path), 'acq_%d.npy' % r))
    assert (got[:2] == want).all()
    assert int(got[3]) == want_ts
    assert got[4] == want_after
    sizes.append(int(got[2]))
  assert sorted(sizes) == [500, 501]                  # each rank scored only its shard
