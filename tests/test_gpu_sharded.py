"""
Multi-rank acquisitions on the real device path (-m gpu): two processes (gloo for the 16-byte collective, both
computing on cuda:0 -- the single-GPU test box; NCCL refuses two ranks on one device) run the SAME seeded
acquisition calls with candidate sharding on, and must return exactly the single-process recommendations:
EI / UCB (one posterior, candidate shards), Add-UCB (one collective per group), Thompson sampling (block shards),
multi-objective UCB and TS.  SURVEY.md 8e.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_all(counts=None):
  """ The acquisition calls under test; returns {name: recommended point}. """
  sys.path.insert(0, ROOT)
  from dragonfly_b200 import kernel, gp_core, domains, synth_data
  from dragonfly_b200 import gpb_acquisitions as A
  from dragonfly_b200 import multiobjective_gpb_acquisitions as M
  out = {}
  w = synth_data.make_workload('c2_hartmann6_matern_ucb', n_train=400, n_cand=10)
  k = w['kernel']
  X, Y = w['X'], w['Y']
  gp = gp_core.GP(X, Y, kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                  gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  if counts is not None:
    from contextlib import contextmanager
    orig = gp._fused_session
    @contextmanager
    def counting(*a, **kw):
      with orig(*a, **kw) as sess:
        inner = sess.score
        def score(pts, **k2):
          counts.append(len(pts))
          return inner(pts, **k2)
        sess.score = score
        yield sess
    gp._fused_session = counting
  dom = domains.EuclideanDomain([[0, 1]] * 6)

  def anc(name, max_evals, in_progress=()):
    return Namespace(curr_acq=name, max_evals=max_evals, t=400, domain=dom, curr_max_val=float(Y.max()),
                     eval_points_in_progress=list(in_progress), acq_opt_method='rand', handle_parallel='halluc',
                     mf_strategy=None, is_mf=False, domain_bounds=np.array(dom.bounds),
                     obj_weights=np.array([0.7, 0.3]), reference_point=[0.0, -1.0])
  for name in ['ei', 'ucb', 'pi', 'ttei']:
    np.random.seed(21)
    out[name] = getattr(A.asy, name)(gp, anc(name, 3001))
  Xh = list(np.random.RandomState(8).random_sample((2, 6)))
  np.random.seed(22)
  out['ucb_halluc'] = A.asy.ucb(gp, anc('ucb', 2000, in_progress=Xh))
  np.random.seed(23)
  out['ts'] = A.asy.ts(gp, anc('ts', 9000))               # 3 blocks of 4096: ranks get 2 + 1
  vals, idxs = gp.draw_samples_argmax(5, np.random.RandomState(31).random_sample((9000, 6)), seed=77)
  out['ts_device_rng_idx'] = idxs.astype(np.float64)
  out['ts_device_rng_val'] = vals
  # additive GP, Add-UCB
  rs = np.random.RandomState(0)
  Xa = rs.random_sample((150, 8))
  Ya = np.sin(3 * Xa[:, 0]) + Xa[:, 3] * Xa[:, 4] - (Xa[:, 6] - 0.5) ** 2
  groups = [[0, 1, 2], [3, 4, 5], [6, 7]]
  add_k = kernel.AdditiveKernel(float(Ya.var()) / 3.0,
                                [kernel.MaternKernel(len(g), 2.5, 1.0, [0.5] * len(g)) for g in groups], groups)
  gpa = gp_core.GP(Xa, Ya, add_k, gp_core.ConstantMean(float(np.median(Ya))), 0.01 * float(Ya.var()))
  doma = domains.EuclideanDomain([[0, 1]] * 8)
  anca = Namespace(curr_acq='add_ucb', max_evals=3000, t=150, domain=doma, curr_max_val=float(Ya.max()),
                   eval_points_in_progress=[], acq_opt_method='rand', handle_parallel='halluc', mf_strategy=None,
                   is_mf=False, domain_bounds=np.array(doma.bounds))
  np.random.seed(24)
  out['add_ucb'] = A.asy.add_ucb(gpa, anca)
  # two objectives
  Y2 = -np.sum((X - 0.4) ** 2, axis=1)
  gp2 = gp_core.GP(X, Y2, kernel.SEKernel(6, float(Y2.var()), [0.35] * 6),
                   gp_core.ConstantMean(float(np.median(Y2))), 0.01 * float(Y2.var()))
  for name, evals in [('lin_ucb', 2500), ('tch_ucb', 2500), ('lin_ts', 9000), ('tch_ts', 5000)]:
    np.random.seed(25)
    out['moo_' + name] = getattr(M.asy, name)([gp, gp2], anc(name, evals))
  return out


def _worker(rank, world, port, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(0)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  counts = []
  res = run_all(counts)
  np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), counts=np.array(counts), **res)
  dist.barrier()
  dist.destroy_process_group()


def test_two_ranks_recommend_what_one_process_recommends(tmp_path):
  port = 35500 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  want = run_all()
  got = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(2)]
  for name, pt in want.items():
    for r in range(2):
      assert (got[r][name] == pt).all(), (name, r, got[r][name], pt)
  # each rank scored about half of every candidate set (the first call: EI over 3001 candidates)
  assert sorted([int(got[0]['counts'][0]), int(got[1]['counts'][0])]) == [1500, 1501]
