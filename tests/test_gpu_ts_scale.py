"""
Thompson sampling at scale (-m gpu): device-side counter-based normals (dfb_fill_rng) and the running per-draw
arg-max (dfb_ts_argmax) behind GP.draw_samples_argmax (BASELINE config 5; asy_ts, gpb_acquisitions.py:119-127;
draw_samples, gp_core.py:250-254).  The generator is pinned bit for bit against oracle/philox.py (itself pinned on
the published Philox known-answer vectors); the draws are pinned against the oracle's draw_gaussian_samples fed
with the very normals the device generated.
"""
from argparse import Namespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def B():
  import torch
  assert torch.cuda.is_available()
  from dragonfly_b200 import kernel, gp_core, device, _lib, synth_data
  from oracle import gp_oracle as O, philox
  return Namespace(kernel=kernel, gp_core=gp_core, device=device, lib=_lib, synth=synth_data, O=O, P=philox,
                   torch=torch)


@pytest.fixture(scope='module')
def small_gp(B):
  w = B.synth.make_workload('c2_hartmann6_matern_ucb', n_train=300, n_cand=5000)
  k = w['kernel']
  gp = B.gp_core.GP(w['X'], w['Y'], B.kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                    B.gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  ogp = B.O.OGP(w['X'], w['Y'], B.O.OMaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                lambda x: np.array([w['mean_const']] * len(x)), w['noise_var'])
  return w, gp, ogp


def test_fill_rng_is_the_published_philox(B, small_gp):
  _, gp, _ = small_gp
  post = gp._post
  for seed, col0 in [(0, 0), (12345678901234567, 2 ** 32 - 300), (7, 10 ** 12)]:
    u = post.fill_rng(seed, col0, 5, 700, what=B.lib.DFB_RNG_UNIFORM).cpu().numpy()
    assert (u == B.P.fill(seed, col0, 5, 700, what=1)).all()          # integer pipeline: exact
    z = post.fill_rng(seed, col0, 5, 700).cpu().numpy()
    np.testing.assert_allclose(z, B.P.fill(seed, col0, 5, 700), rtol=1e-12, atol=1e-13)   # libm vs CUDA log / cos
  # a candidate's normals do not depend on how the columns are cut into blocks
  whole = post.fill_rng(9, 0, 3, 1000).cpu().numpy()
  parts = np.concatenate([post.fill_rng(9, 0, 3, 400).cpu().numpy(), post.fill_rng(9, 400, 3, 600).cpu().numpy()], axis=1)
  assert (whole == parts).all()
  big = post.fill_rng(1, 0, 64, 20000).cpu().numpy()
  assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3 and np.isfinite(big).all()
  assert abs(np.corrcoef(big[0], big[1])[0, 1]) < 0.03 and abs(np.corrcoef(big[0, :-1], big[0, 1:])[0, 1]) < 0.03
  with pytest.raises(B.lib.DfbError):
    post.fill_rng(1, 0, 2, 10, what=5)


def test_ts_argmax_order(B, small_gp):
  _, gp, _ = small_gp
  post, t = gp._post, B.torch
  smp = t.zeros((3, 1000), dtype=t.float64, device='cuda')
  smp[0, [17, 400]] = 2.0                         # tie -> first index
  smp[1, 600] = float('nan'); smp[1, 10] = 9.0    # NaN counts as the maximum
  smp[2, :] = -5.0; smp[2, 999] = -4.0
  best = t.empty(3, dtype=t.float64, device='cuda'); idx = t.empty(3, dtype=t.int64, device='cuda')
  post.ts_argmax(smp, 5000, best, idx, reset=True)
  assert idx.cpu().tolist() == [5017, 5600, 5999]
  more = t.full((3, 50), -1.0, dtype=t.float64, device='cuda')
  more[0, 3] = 2.0                                # equal to the running best: the earlier (lower) index stays
  more[2, 7] = 100.0
  post.ts_argmax(more, 100, best, idx, reset=False)
  assert idx.cpu().tolist() == [103, 5600, 107]   # 103 < 5017: np.argmax order is by global index
  assert np.isnan(best.cpu().numpy()[1]) and best.cpu().numpy()[2] == 100.0


def test_draw_samples_argmax_equals_oracle_on_the_same_normals(B, small_gp):
  w, gp, ogp = small_gp
  C, S, seed = w['candidates'], 6, 5
  vals, idxs = gp.draw_samples_argmax(S, C, seed=seed)
  assert vals.shape == (S,) and idxs.shape == (S,)
  want_v, want_i = np.full(S, -np.inf), np.full(S, -1)
  for lo in range(0, len(C), 4096):
    hi = min(len(C), lo + 4096)
    Ut = gp._post.fill_rng(seed, lo, S, hi - lo).cpu().numpy()               # the normals the device used
    smp = ogp.draw_samples_with_normals(C[lo:hi], np.ascontiguousarray(Ut.T))   # oracle: (S, m)
    info, dev_smp, _ = gp._post.ts_draws(C[lo:hi], Ut, mean_const=gp._mean_const)
    if info == 0:                                   # (a block that needs the jitter ladder is covered by the arg-max check)
      np.testing.assert_allclose(dev_smp.cpu().numpy(), smp, rtol=0, atol=2e-6)
    for s in range(S):
      j = int(np.argmax(smp[s]))
      if smp[s, j] > want_v[s]:
        want_v[s], want_i[s] = smp[s, j], lo + j
  assert (idxs == want_i).all()
  np.testing.assert_allclose(vals, want_v, rtol=0, atol=2e-6)
  # reproducible, seed-dependent, and the draws scatter like the posterior says
  v2, i2 = gp.draw_samples_argmax(S, C, seed=seed)
  assert (i2 == idxs).all() and (v2 == vals).all()
  _, i3 = gp.draw_samples_argmax(S, C, seed=seed + 1)
  assert (i3 != idxs).any()
  # hallucinated variant runs through the in-place extension and leaves the posterior intact
  mu0, sd0 = gp.eval(C[:64], 'std')
  Xh = list(np.random.RandomState(2).random_sample((2, 6)))
  vh, ih = gp.draw_samples_argmax(S, C, seed=seed, X_halluc=Xh)
  assert ih.shape == (S,) and np.isfinite(vh).all()
  mu1, sd1 = gp.eval(C[:64], 'std')
  assert (mu0 == mu1).all() and (sd0 == sd1).all()


def test_config5_geometry_runs_and_is_fast(B):
  """ Park1-20D, N = 5000, 64 draws x 3 blocks: finite, indices in range, wall-clock printed. """
  import time
  w = B.synth.make_workload('c5_park1_20_ts', n_cand=3 * 4096)
  k = w['kernel']
  gp = B.gp_core.GP(w['X'], w['Y'], B.kernel.MaternKernel(20, 2.5, k['scale'], k['dim_bandwidths']),
                    B.gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  Cd = B.torch.from_numpy(w['candidates']).cuda()
  gp.draw_samples_argmax(64, Cd[:4096], seed=1)
  B.torch.cuda.synchronize(); t0 = time.perf_counter()
  vals, idxs = gp.draw_samples_argmax(64, Cd, seed=1)
  B.torch.cuda.synchronize(); dt = time.perf_counter() - t0
  print('C5 geometry: 64 draws x %d candidates in %.1f ms (%.0f candidates/s)' % (len(Cd), 1e3 * dt, len(Cd) / dt))
  assert np.isfinite(vals).all() and (idxs >= 0).all() and (idxs < len(Cd)).all()
  assert len(set(idxs.tolist())) > 1
