"""
GP.eval of a handful of points (-m gpu): the row-streaming kernel behind dfb_eval for m <= 32 (the one-point
objective of the sequential maximisers, TTEI's reference arm, BOCA's fidelity scan; gp_core.py:165-190) against
the tile kernels on the same points and against the oracle / reference golden.
"""
from argparse import Namespace

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def B():
  import torch
  assert torch.cuda.is_available()
  from dragonfly_b200 import kernel, gp_core, device, _lib, synth_data
  from oracle import gp_oracle as O
  return Namespace(kernel=kernel, gp_core=gp_core, device=device, lib=_lib, synth=synth_data, O=O, torch=torch)


@pytest.mark.parametrize('n', [50, 300, 1500])
def test_small_batches_match_tile_kernels_and_oracle(B, n):
  w = B.synth.make_workload('c2_hartmann6_matern_ucb', n_train=n, n_cand=400)
  k = w['kernel']
  gp = B.gp_core.GP(w['X'], w['Y'], B.kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                    B.gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  ogp = B.O.OGP(w['X'], w['Y'], B.O.OMaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                lambda x: np.array([w['mean_const']] * len(x)), w['noise_var'])
  C = w['candidates']
  mu_big, sd_big = gp.eval(C, 'std')                       # 400 points: tile kernels
  mu_o, var_o = B.O.eval_std_diag(ogp, C[:32])
  for m in [1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 28, 32]:
    mu, sd = gp.eval(C[:m], 'std')
    np.testing.assert_array_equal(mu, mu_big[:m])          # the mean comes from the same K_* kernel
    np.testing.assert_allclose(sd ** 2, sd_big[:m] ** 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sd ** 2, var_o[:m], rtol=0, atol=1e-8)
    np.testing.assert_allclose(mu, mu_o[:m], rtol=0, atol=1e-10)
  # 33 points are back on the tile path: bit-identical to the big batch
  mu33, sd33 = gp.eval(C[:33], 'std')
  assert (sd33 == sd_big[:33]).all()
  # a point's result does not depend on what else is in a small batch (the level-wide PDOO batches rely on it)
  mu_a, sd_a = gp.eval(C[5:6], 'std'); mu_b, sd_b = gp.eval(C[:29], 'std')
  assert sd_a[0] == sd_b[5] and mu_a[0] == mu_b[5]
  # the switch
  gp._post.set_option('small_eval', 0)
  mu1, sd1 = gp.eval(C[:1], 'std')
  assert (sd1 == sd_big[:1]).all()
  gp._post.set_option('small_eval', 1)
  # device-tensor candidates take the same path
  mu_d, sd_d = gp.eval(B.torch.from_numpy(C[:3]).cuda(), 'std')
  mu_h, sd_h = gp.eval(C[:3], 'std')
  assert (sd_d.cpu().numpy() == sd_h).all() and (mu_d.cpu().numpy() == mu_h).all()


def test_single_points_against_the_reference_golden(B):
  g = load_golden('c1_se')
  gp = B.gp_core.GP(g['X'], g['Y'], B.kernel.SEKernel(2, float(g['scale']), g['bws']),
                    B.gp_core.ConstantMean(float(g['mean_const'])), float(g['noise_var']))
  for i in [0, 1, 17, 555]:
    mu, sd = gp.eval(g['C'][i:i + 1], 'std')
    assert abs(mu[0] - g['mu'][i]) <= 1e-10 and abs(sd[0] ** 2 - g['sd'][i] ** 2) <= 1e-8
  mu, sd = gp.eval(g['C'][:11], 'std')
  np.testing.assert_allclose(mu, g['mu'][:11], rtol=0, atol=1e-10)
  np.testing.assert_allclose(sd ** 2, g['sd'][:11] ** 2, rtol=0, atol=1e-8)


def test_one_point_latency_at_n5000(B):
  """ Prints the per-call time of gp.eval(1 point) at the metric's N with and without the row-streaming kernel. """
  import time
  w = B.synth.make_workload('headline_hartmann6_matern_ei', n_cand=64)
  k = w['kernel']
  gp = B.gp_core.GP(w['X'], w['Y'], B.kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                    B.gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  x = w['candidates'][:1]
  res = {}
  for opt in (1, 0):
    gp._post.set_option('small_eval', opt)
    for _ in range(5):
      gp.eval(x, 'std')
    B.torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
      mu, sd = gp.eval(x, 'std')
    B.torch.cuda.synchronize()
    res[opt] = (1e3 * (time.perf_counter() - t0) / 50, float(sd[0]))
  print('gp.eval(1 point, std) at N=5000: row-streaming %.3f ms, tile kernels %.3f ms' % (res[1][0], res[0][0]))
  assert abs(res[1][1] ** 2 - res[0][1] ** 2) <= 1e-12


@pytest.mark.parametrize('acq', ['ei', 'ucb'])
def test_pdoo_maximiser_end_to_end(B, acq):
  """ acq_opt_method='pdoo' (what 'direct' falls back to without the Fortran extension): the batched PDOO of
      dragonfly_b200/doo.py driving the device-backed acquisition a whole round of all its passes per call, against the reference's
      asy_ei / asy_ucb recommendation (tests/golden/pdoo.npz; its doo.py evaluates one point per call). """
  from dragonfly_b200 import gpb_acquisitions as A, domains, doo
  g = load_golden('pdoo')
  w = B.synth.make_workload('c1_branin_se_ei', n_cand=10)
  k = w['kernel']
  gp = B.gp_core.GP(w['X'], w['Y'], B.kernel.SEKernel(2, k['scale'], k['dim_bandwidths']),
                    B.gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  anc = Namespace(curr_acq=acq, max_evals=150, t=50, domain=domains.EuclideanDomain([[0, 1]] * 2),
                  curr_max_val=float(w['Y'].max()), eval_points_in_progress=[], acq_opt_method='pdoo',
                  handle_parallel='halluc', mf_strategy=None, is_mf=False)
  pt = getattr(A.asy, acq)(gp, anc)
  s = doo.pdoo_maximise.last_search
  assert s.num_device_calls < 0.25 * len(s.query_vals) + 10 and s.num_prefetched > 0
  want = g['e2e_%s_pdoo_point' % acq]
  if not (np.asarray(pt) == want).all():
    # a tree search may legitimately branch differently on a 1e-12 difference of two near-equal bounds: then
    # the recommendation must at least be as good as the reference's under the oracle's acquisition
    ogp = B.O.OGP(w['X'], w['Y'], B.O.OSEKernel(2, k['scale'], k['dim_bandwidths']),
                  lambda x: np.array([w['mean_const']] * len(x)), w['noise_var'])
    def score(x):
      mu, sd = ogp.eval(np.asarray(x).reshape(1, -1), 'std')
      if acq == 'ei':
        return float(B.O.acq_ei(mu, sd, float(w['Y'].max()))[0])
      return float(B.O.acq_ucb(mu, sd, B.O.ucb_beta_th(2, 50))[0])
    assert score(pt) >= score(want) - 1e-9, (pt, want)
  # 'direct' without a Fortran DIRECT build is the same search (oper_utils.py:121-137)
  anc.acq_opt_method = 'direct'
  if not A._reference_fortran_direct_available():
    assert (getattr(A.asy, acq)(gp, anc) == pt).all()
