"""
Pins the NumPy oracle (oracle/gp_oracle.py) against (1) the reference's own known-answer vectors
and (2) outputs of the unmodified reference recorded in tests/golden/*.npz by make_golden.py.
CPU only.  Tolerances: the oracle follows the reference's operation order, so agreement is at the
BLAS-summation-order level (<= 1e-12 relative); arg-max indices are exact.
"""
import numpy as np
import pytest

from oracle import gp_oracle as O
from conftest import load_golden


def const_mean(c):
  return lambda x: np.array([c] * len(x))


def close(a, b, rtol=1e-12, atol=1e-13):
  np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


# ---- (1) the reference's own known-answer tests ------------------------------------------------
def test_dist_squared_known_answer_exact():
  """ dragonfly/utils/unittest_general_utils.py:26-35 -- exact equality. """
  X1 = np.array([[1, 2, 3], [1, 2, 4], [2, 3, 4.5]])
  X2 = np.array([[1, 2, 4], [1, 2, 5], [2, 3, 5]])
  true = np.array([[1, 4, 6], [0, 1, 3], [2.25, 2.25, 0.25]])
  assert (O.dist_squared(X1, X2) == true).all()
  g = load_golden('known_answers')
  assert (O.dist_squared(g['ds_X1'], g['ds_X2']) == g['ds_true']).all()


def test_se_kernel_known_answer():
  """ dragonfly/gp/unittest_kernel.py:37-47,82-91 -- closed form, tol 1e-10. """
  d1 = np.array([[1, 2], [3, 4.5]]); d2 = np.array([[1, 2], [3, 4]])
  kern = O.OSEKernel(2, 2, [0.1, 1])
  t11 = 2 * np.array([[1, np.exp(-406.25 / 2)], [np.exp(-406.25 / 2), 1]])
  t22 = 2 * np.array([[1, np.exp(-404 / 2)], [np.exp(-404 / 2), 1]])
  t12 = 2 * np.array([[1, np.exp(-404 / 2)], [np.exp(-406.25 / 2), np.exp(-0.25 / 2)]])
  assert np.linalg.norm(t11 - kern(d1)) < 1e-10
  assert np.linalg.norm(t22 - kern(d2)) < 1e-10
  assert np.linalg.norm(t12 - kern(d1, d2)) < 1e-10
  g = load_golden('known_answers')
  close(kern(d1), g['se_11']); close(kern(d2), g['se_22']); close(kern(d1, d2), g['se_12'])


@pytest.mark.parametrize('nu', [0.5, 1.5, 2.5])
def test_matern_kernel_known_answer(nu):
  """ dragonfly/gp/unittest_kernel.py:49-54,93-124 -- closed forms, tol 1e-10. """
  d1 = np.array([[1, 2], [3, 4.5]]); d2 = np.array([[1, 2], [3, 4]])
  sd11 = np.array([[0, np.sqrt(406.25)], [np.sqrt(406.25), 0]])
  sd12 = np.array([[0, np.sqrt(404)], [np.sqrt(406.25), np.sqrt(0.25)]])

  def closed(dist):
    if nu == 0.5:
      return 2.1 * np.exp(-dist)
    if nu == 1.5:
      return 2.1 * np.exp(-np.sqrt(3) * dist) * (1 + np.sqrt(3) * dist)
    return 2.1 * np.exp(-np.sqrt(5) * dist) * (1 + np.sqrt(5) * dist + (5 / 3.0) * dist ** 2)
  kern = O.OMaternKernel(2, nu, 2.1, [0.1, 1])
  assert np.linalg.norm(closed(sd11) - (1.0 / kern.norm_constant) * kern(d1)) < 1e-10
  assert np.linalg.norm(closed(sd12) - (1.0 / kern.norm_constant) * kern(d1, d2)) < 1e-10
  g = load_golden('known_answers')
  tag = str(nu).replace('.', 'p')
  assert kern.norm_constant == float(g['matern_%s_norm_constant' % tag])
  close(kern(d1), g['matern_%s_11' % tag]); close(kern(d1, d2), g['matern_%s_12' % tag])


def test_empty_inputs_follow_reference():
  """ kernel.py:79-81 (zeros for empty sides), general_utils.py:210-211, :173-174. """
  kern = O.OSEKernel(2, 1.0, [1.0, 1.0])
  assert kern(np.zeros((0, 2)), np.zeros((3, 2))).shape == (0, 3)
  assert kern(np.zeros((3, 2)), np.zeros((0, 2))).shape == (3, 0)
  assert O.solve_lower_triangular(np.zeros((0, 0)), np.zeros((0, 4))).shape == (0, 4)
  L, p = O.stable_cholesky(np.zeros((0, 0)))
  assert L.size == 0 and p is None


# ---- (2) reference outputs recorded by make_golden.py -------------------------------------------
def _gp_c1():
  g = load_golden('c1_se')
  kern = O.OSEKernel(2, float(g['scale']), g['bws'])
  gp = O.OGP(g['X'], g['Y'], kern, const_mean(float(g['mean_const'])), float(g['noise_var']))
  return g, gp


def test_c1_posterior_state():
  g, gp = _gp_c1()
  close(gp.K_trtr_wo_noise, g['K'])
  close(gp.L, g['L'], rtol=1e-11, atol=1e-12)
  close(gp.alpha, g['alpha'], rtol=1e-10, atol=1e-12)
  close(gp.compute_log_marginal_likelihood(), g['lml'], rtol=1e-12)


def test_c1_eval_and_acquisitions():
  g, gp = _gp_c1()
  mu, sd = gp.eval(g['C'], 'std')
  close(mu, g['mu'], rtol=0, atol=1e-10)
  close(sd ** 2, g['sd'] ** 2, rtol=0, atol=1e-8)
  mu0, none = gp.eval(g['C'], 'none')
  assert none is None
  close(mu0, g['mu_none'], rtol=0, atol=1e-10)
  best = float(g['curr_best'])
  ei = O.acq_ei(mu, sd, best); pi = O.acq_pi(mu, sd, best); ucb = O.acq_ucb(mu, sd, float(g['beta']))
  close(O.ucb_beta_th(2, int(g['t'])), g['beta'], rtol=1e-15)
  close(ei, g['ei'], rtol=1e-9, atol=1e-12); close(pi, g['pi'], rtol=1e-9, atol=1e-12)
  close(ucb, g['ucb'], rtol=0, atol=1e-9)
  ri = int(g['ttei_ref_idx'])
  ttei = O.acq_ttei(mu, sd, mu[ri], sd[ri])
  close(ttei, g['ttei'], rtol=1e-9, atol=1e-12)
  assert O.np_argmax_first(ei) == int(g['argmax_ei'])
  assert O.np_argmax_first(pi) == int(g['argmax_pi'])
  assert O.np_argmax_first(ucb) == int(g['argmax_ucb'])
  assert O.np_argmax_first(ttei) == int(g['argmax_ttei'])
  with pytest.raises(ValueError):
    gp.eval(g['C'][:4], 'bogus')


def test_c1_chunked_driver_matches_unchunked():
  g, gp = _gp_c1()
  val, idx, scores = O.chunked_scores(gp, g['C'], 'ei', chunk=512, curr_best=float(g['curr_best']))
  close(scores, g['ei'], rtol=1e-9, atol=1e-12)
  assert idx == int(g['argmax_ei'])


def test_c1_end_to_end_random_maximise():
  """ asy_ei/ucb/pi through maximise_acquisition + random_maximise with the same uniforms. """
  g, gp = _gp_c1()
  pts = O.map_to_bounds(g['e2e_U'], [[0, 1]] * 2)
  best = float(g['curr_best'])
  for name, fn in [('ei', lambda x: O.acq_ei(*gp.eval(x, 'std'), best)),
                   ('pi', lambda x: O.acq_pi(*gp.eval(x, 'std'), best)),
                   ('ucb', lambda x: O.acq_ucb(*gp.eval(x, 'std'), float(g['beta'])))]:
    _, pt, _, _ = O.random_maximise_on_points(fn, pts)
    assert (pt == g['e2e_%s_point' % name]).all()


def test_c1_hallucinated():
  g, gp = _gp_c1()
  mu_h, sd_h = gp.eval_with_hallucinated_observations(g['C'][:800], list(g['Xh']), 'std')
  close(mu_h, g['mu_h'], rtol=0, atol=1e-10)
  close(sd_h ** 2, g['sd_h'] ** 2, rtol=0, atol=1e-8)
  mu_d, var_d = O.eval_std_diag(gp, g['C'][:800], X_halluc=list(g['Xh']))
  close(var_d, g['sd_h'] ** 2, rtol=0, atol=1e-8)


@pytest.mark.parametrize('nu', [0.5, 1.5, 2.5])
def test_matern_h6(nu):
  g = load_golden('matern_h6')
  tag = str(nu).replace('.', 'p')
  kern = O.OMaternKernel(6, nu, float(g['scale']), g['bws'])
  gp = O.OGP(g['X'], g['Y'], kern, const_mean(float(g['mean_const'])), float(g['noise_var']))
  close(gp.alpha, g['alpha_' + tag], rtol=1e-9, atol=1e-11)
  close(np.diag(gp.L), g['Ldiag_' + tag], rtol=1e-11)
  close(gp.L[::7, ::5], g['Lsub_' + tag], rtol=1e-10, atol=1e-12)
  close(gp.K_trtr_wo_noise[::7, ::5], g['Ksub_' + tag])
  close(gp.compute_log_marginal_likelihood(), g['lml_' + tag], rtol=1e-12)
  close(kern(g['C'][:64], g['X']), g['Kstar_sub_' + tag])
  mu, sd = gp.eval(g['C'], 'std')
  close(mu, g['mu_' + tag], rtol=0, atol=1e-10)
  close(sd ** 2, g['sd_' + tag] ** 2, rtol=0, atol=1e-8)
  ucb = O.acq_ucb(mu, sd, O.ucb_beta_th(6, int(g['t'])))
  ei = O.acq_ei(mu, sd, float(g['curr_best']))
  assert O.np_argmax_first(ucb) == int(g['argmax_ucb_' + tag])
  assert O.np_argmax_first(ei) == int(g['argmax_ei_' + tag])
  mu_d, var_d = O.eval_std_diag(gp, g['C'])
  close(mu_d, g['mu_' + tag], rtol=0, atol=1e-10)
  close(var_d, g['sd_' + tag] ** 2, rtol=0, atol=1e-8)


def _additive_gp():
  g = load_golden('additive')
  groups = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
  sub = [O.OMaternKernel(4, 2.5, 1.0, [0.5] * 4), O.OSEKernel(4, 1.0, [0.4, 0.5, 0.6, 0.7]),
         O.OMaternKernel(2, 1.5, 1.0, [0.3, 0.45])]
  kern = O.OAdditiveKernel(float(g['scale']), sub, groups)
  gp = O.OGP(g['X'], g['Y'], kern, const_mean(float(g['mean_const'])), float(g['noise_var']))
  return g, gp, kern


def test_additive_and_add_ucb():
  g, gp, kern = _additive_gp()
  close(gp.alpha, g['alpha'], rtol=1e-9, atol=1e-11)
  close(gp.K_trtr_wo_noise[::5, ::3], g['Ksub'])
  close(gp.compute_log_marginal_likelihood(), g['lml'], rtol=1e-12)
  mu, sd = gp.eval(g['C'], 'std')
  close(mu, g['mu'], rtol=0, atol=1e-10); close(sd ** 2, g['sd'] ** 2, rtol=0, atol=1e-8)
  for j in range(3):
    score, mu_j, sd_j = O.add_ucb_group_scores(gp, kern, j, g['Cj_%d' % j], int(g['t']))
    close(mu_j, g['mu_j_%d' % j], rtol=0, atol=1e-10)
    close(sd_j ** 2, g['sd_j_%d' % j] ** 2, rtol=0, atol=1e-8)
    close(score, g['score_j_%d' % j], rtol=0, atol=1e-9)
    assert O.np_argmax_first(score) == int(g['argmax_j_%d' % j])
  pts = [g['e2e_U_%d' % j] for j in range(3)]
  ret, _ = O.add_ucb_on_points(gp, kern, pts, int(g['t']))
  assert (ret == g['e2e_point']).all()


def test_mf_product_kernel_and_fidel_slice():
  g = load_golden('mf')
  kF = O.OSEKernel(1, 1.0, [0.7]); kD = O.OMaternKernel(4, 2.5, 1.0, [0.4] * 4)
  kern = O.OCoordinateProductKernel(5, float(g['scale']), [kF, kD], [[0], [1, 2, 3, 4]])
  ZX = np.concatenate((g['Z'], g['Xd']), axis=1)
  gp = O.OGP(ZX, g['Y'], kern, const_mean(float(g['mean_const'])), float(g['noise_var']))
  close(gp.alpha, g['alpha'], rtol=1e-9, atol=1e-11)
  close(gp.compute_log_marginal_likelihood(), g['lml'], rtol=1e-12)
  mu, sd = gp.eval(np.concatenate((g['Cz'], g['Cx']), axis=1), 'std')
  close(mu, g['mu'], rtol=0, atol=1e-10); close(sd ** 2, g['sd'] ** 2, rtol=0, atol=1e-8)
  mu_f, sd_f = gp.eval(O.mf_zx(g['f2o'], g['Cx']), 'std')
  close(mu_f, g['mu_f'], rtol=0, atol=1e-10); close(sd_f ** 2, g['sd_f'] ** 2, rtol=0, atol=1e-8)
  ucb = O.acq_ucb(mu_f, sd_f, O.ucb_beta_th(4, int(g['t'])))
  close(ucb, g['ucb_f'], rtol=0, atol=1e-9)
  assert O.np_argmax_first(ucb) == int(g['argmax_ucb_f'])


def test_thompson_draws_with_supplied_normals():
  g = load_golden('ts')
  kern = O.OMaternKernel(20, 2.5, float(g['scale']), g['bws'])
  gp = O.OGP(g['X'], g['Y'], kern, const_mean(float(g['mean_const'])), float(g['noise_var']))
  samples = gp.draw_samples_with_normals(g['C'], g['U'])
  close(samples, g['samples'], rtol=0, atol=1e-7)
  assert (samples.argmax(axis=1) == g['argmax']).all()


def test_jitter_ladder():
  g = load_golden('jitter')
  kern = O.OSEKernel(3, float(g['scale']), g['bws'])
  gp = O.OGP(g['X'], g['Y'], kern, const_mean(0.0), 0.0)
  assert int(g['power']) != -999
  assert gp.jitter_power == int(g['power'])
  close(gp.L, g['L'], rtol=1e-9, atol=1e-12)


def test_lml_grid_and_sampling_weights():
  g = load_golden('lml_grid')
  lmls = []
  for hp in g['hps']:
    kern = O.OMaternKernel(6, 2.5, np.exp(hp[1]), np.exp(hp[2:]))
    gp = O.OGP(g['X'], g['Y'], kern, const_mean(float(g['mean_const'])), np.exp(hp[0]))
    lmls.append(gp.compute_log_marginal_likelihood())
  close(lmls, g['lmls'], rtol=1e-11)
  w = O.rand_exp_sampling_probs(lmls)
  assert abs(w.sum() - 1) < 1e-12 and w.argmax() == np.argmax(g['lmls'])


# ---- multi-objective scalarisations and add_data ----------------------------------------
def _moo_oracle_gps(g):
  gp1 = O.OGP(g['X'], g['Y1'], O.OMaternKernel(6, 2.5, float(g['scale1']), g['bws1']),
              const_mean(float(g['mean1'])), float(g['noise1']))
  gp2 = O.OGP(g['X'], g['Y2'], O.OSEKernel(6, float(g['scale2']), g['bws2']),
              const_mean(float(g['mean2'])), float(g['noise2']))
  return [gp1, gp2]


def test_moo_scalarisations_match_reference():
  """ multiobjective_gpb_acquisitions.py:79-107 on the reference's own (mu, sigma). """
  g = load_golden('moo')
  gps = _moo_oracle_gps(g)
  beta = O.moo_ucb_beta_th(6, int(g['t']))
  assert beta == float(g['beta'])
  mus, sds = zip(*[gp.eval(g['C'], 'std') for gp in gps])
  lin = O.moo_lin_ucb(mus, sds, g['weights'], beta)
  tch = O.moo_tch_ucb(mus, sds, g['weights'], g['refs'], beta)
  close(lin, g['lin_ucb_scores'], rtol=1e-10, atol=1e-11)
  close(tch, g['tch_ucb_scores'], rtol=1e-10, atol=1e-11)
  assert O.np_argmax_first(lin) == int(np.argmax(g['lin_ucb_scores']))
  assert O.np_argmax_first(tch) == int(np.argmax(g['tch_ucb_scores']))


def test_moo_end_to_end_recommendations():
  """ The reference's recommendations under np.random.seed(9): candidates from np.random.random, then (TS)
      one np.random.normal draw per objective, scalarise, arg-max. """
  g = load_golden('moo')
  gps = _moo_oracle_gps(g)
  beta = float(g['beta'])
  bounds = [[0, 1]] * 6
  for name in ['lin_ucb', 'tch_ucb']:
    np.random.seed(9)
    pts = O.map_to_bounds(np.random.random((1500, 6)), bounds)
    mus, sds = zip(*[gp.eval(pts, 'std') for gp in gps])
    sc = (O.moo_lin_ucb(mus, sds, g['weights'], beta) if name == 'lin_ucb'
          else O.moo_tch_ucb(mus, sds, g['weights'], g['refs'], beta))
    assert (pts[O.np_argmax_first(sc)] == g['e2e_%s_point' % name]).all()
  for name, halluc in [('lin_ts', None), ('tch_ts', None), ('lin_ts_halluc', g['Xh'])]:
    np.random.seed(9)
    pts = O.map_to_bounds(np.random.random((300, 6)), bounds)
    samples = []
    for gp in gps:
      U = np.random.normal(size=(300, 1))
      samples.append(gp.draw_samples_with_normals(pts, U, None if halluc is None else list(halluc)).ravel())
    sc = (O.moo_tch_vals(samples, g['weights'], g['refs']) if name == 'tch_ts'
          else O.moo_lin_vals(samples, g['weights']))
    assert (pts[O.np_argmax_first(sc)] == g['e2e_%s_point' % name]).all()


def test_add_data_matches_reference():
  """ gp_core.py:135-146: the reference after add_data_multiple + 2 x add_data_single. """
  g = load_golden('incremental')
  n0 = int(g['n0'])
  gp = O.OGP(g['X'][:n0], g['Y'][:n0], O.OMaternKernel(6, 2.5, float(g['scale']), g['bws']),
             const_mean(float(g['mean_const'])), float(g['noise_var']))
  gp.add_data_multiple(list(g['X'][n0:n0 + 4]), list(g['Y'][n0:n0 + 4]))
  gp.add_data_multiple([g['X'][n0 + 4]], [g['Y'][n0 + 4]])
  gp.add_data_multiple([g['X'][n0 + 5]], [g['Y'][n0 + 5]])
  close(gp.L, g['L'], rtol=1e-10, atol=1e-12)
  close(gp.alpha, g['alpha'], rtol=1e-8, atol=1e-9)
  close(gp.compute_log_marginal_likelihood(), g['lml'], rtol=1e-12)
  mu, sd = gp.eval(g['C'], 'std')
  close(mu, g['mu'], atol=1e-11)
  close(sd ** 2, g['sd'] ** 2, atol=1e-11)


def test_philox_known_answers():
  """ oracle/philox.py against the published Random123 known-answer vectors for Philox4x32-10 (the generator
      behind dfb_fill_rng); moments of the Box-Muller normals. """
  from oracle import philox as P
  u32 = lambda *v: [np.array([x], dtype=np.uint32) for x in v]
  kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
          ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
          ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
           (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
  for ctr, key, want in kats:
    got = P.philox4x32_10(*u32(*ctr), *u32(*key))
    assert tuple(int(g[0]) for g in got) == want
  z = P.fill(11, 0, 32, 8192)
  assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and np.isfinite(z).all()
  u = P.fill(11, 0, 32, 8192, what=1)
  assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.005
  # layout invariance: columns [100, 300) of a wide fill == a fill that starts at column 100
  assert (P.fill(3, 0, 4, 300)[:, 100:] == P.fill(3, 100, 4, 200)).all()


def test_stable_cholesky_and_gaussian_draws_like_the_reference_tests():
  """ dragonfly/utils/unittest_general_utils.py:65-94 restated on the oracle: ||L L^T - M|| < 1e-5 on a random SPD
      matrix; sample mean / covariance of 10^4 draws within the reference's 4-sigma tolerances (seeded here). """
  rs = np.random.RandomState(0)
  M = rs.normal(size=(5, 5)); M = M.dot(M.T)
  L, power = O.stable_cholesky(M)
  assert power is None and np.linalg.norm(L.dot(L.T) - M) < 1e-5
  num_samples, num_pts = 10000, 3
  mu = np.arange(num_pts, dtype=np.float64)
  K = rs.normal(size=(num_pts, num_pts)); K = K.dot(K.T)
  samples = O.draw_gaussian_samples_with_normals(mu, K, rs.normal(size=(num_pts, num_samples)))
  assert samples.shape == (num_samples, num_pts)
  sample_mean = samples.mean(axis=0)
  centred = samples - sample_mean
  sample_covar = centred.T.dot(centred) / num_samples
  assert np.linalg.norm(mu - sample_mean) < 4 * np.linalg.norm(mu) / np.sqrt(num_samples)
  assert np.linalg.norm(K - sample_covar) < 4 * np.linalg.norm(K) / np.sqrt(num_samples)


# ---- LML gradients: the oracle's restatement against the unmodified reference ------------
@pytest.mark.parametrize('name', ['se3', 'm25_6', 'm15_2', 'm05_4'])
def test_lml_gradients_match_reference(name):
  """ gp_core.py:229-240 + kernel.py:202-217, 301-322; goldens: tests/golden/make_golden_grad.py. """
  g = load_golden('grad')
  X, Y, bw = g[name + '_X'], g[name + '_Y'], g[name + '_bw']
  scale, nv, mc, nu = [float(v) for v in g[name + '_meta']]
  d = X.shape[1]
  kern = O.OSEKernel(d, scale, bw) if nu < 0 else O.OMaternKernel(d, nu, scale, bw)
  gp = O.OGP(X, Y, kern, lambda x: np.array([mc] * len(x)), nv)
  params = [('scale', ()), ('noise_var', ()), ('noise_mean', ()), ('same_dim_bandwidths', ())] + \
           [('dim_bandwidths', (j,)) for j in range(d)]
  got = np.array([gp.compute_grad_log_marginal_likelihood(p, *a) for p, a in params])
  np.testing.assert_allclose(got, g[name + '_grads'], rtol=1e-10)
  np.testing.assert_allclose(kern.gradient('same_dim_bandwidths', X[:8], X), g[name + '_G_same'], rtol=1e-12, atol=1e-14)
  np.testing.assert_allclose(kern.gradient('dim_bandwidths', X, X, 1)[:8], g[name + '_G_dim1'], rtol=1e-12, atol=1e-14)


def test_cartesian_product_gp_matches_reference():
  """ The reference's CPGP + CartesianProductKernel (kernel.py:504-538, cartesian_product_gp.py:207-248, default
      'project_first') on Euclidean factors is the coordinate-product kernel on the parts laid side by side: the
      eigen-projection of a PSD matrix moves nothing beyond rounding.  Golden: tests/golden/make_golden_cpgp.py. """
  g = load_golden('cpgp')
  scale, nv, mc = [float(v) for v in g['meta']]
  kern = O.OCoordinateProductKernel(6, scale, [O.OSEKernel(2, 1.0, [0.4, 0.6]), O.OMaternKernel(3, 2.5, 1.0, [0.5, 0.7, 0.9]),
                                               O.OMaternKernel(1, 1.5, 1.0, [0.3])], [[0, 1], [2, 3, 4], [5]])
  gp = O.OGP(g['X'], g['Y'], kern, lambda x: np.array([mc] * len(x)), nv)
  np.testing.assert_allclose(gp.K_trtr_wo_noise[:16], g['K'], rtol=0, atol=1e-12)
  np.testing.assert_allclose(gp.compute_log_marginal_likelihood(), float(g['lml']), rtol=1e-10)
  mu, var = O.eval_std_diag(gp, g['C'])
  np.testing.assert_allclose(mu, g['mu'], rtol=0, atol=1e-10)
  np.testing.assert_allclose(var, g['sd'] ** 2, rtol=0, atol=1e-8)
