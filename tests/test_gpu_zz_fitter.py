"""
hp_grid.fit_gp on the device (-m gpu) against the reference's EuclideanGPFitter.fit_gp (golden fitter.npz:
gp_core.py:783-821 with ml_hp_tune_opt rand / rand_exp_sampling / pdoo under np.random.seed(5)).  The selection
logic and RNG consumption are pinned exactly on the CPU (tests/test_host_logic.py, oracle LMLs); here the LMLs come
from DFB_BUILD_LML_ONLY builds (concurrent lanes), so: identical candidates, LMLs to 1e-10 relative, the same
selected hyper-parameters for the sampling methods, and for the tree search a recommendation at least as good.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def setup():
  import torch
  assert torch.cuda.is_available()
  from dragonfly_b200 import hp_grid
  g = load_golden('fitter')
  layout = hp_grid.EuclideanHPLayout(3, 'matern', mean_func_type=str(g['mean_func_type']),
                                     noise_var_type=str(g['noise_var_type']))
  return hp_grid, g, layout


def test_rand_and_rand_exp_sampling(setup):
  hp_grid, g, layout = setup
  X, Y = g['X'], g['Y']
  np.random.seed(5)
  tag, gp, (cts, dscr) = hp_grid.fit_gp(X, Y, layout, g['rand_bounds'], g['rand_dscr_vals'], method='rand',
                                        max_evals=int(g['rand_max_evals']))
  assert tag == 'fitted_gp'
  assert (np.array(cts) == g['rand_cts']).all() and (np.array(dscr) == g['rand_dscr']).all()
  np.testing.assert_allclose(gp.compute_log_marginal_likelihood(), float(g['rand_lml']), rtol=1e-10)
  mu, sd = gp.eval(g['rand_C'], 'std')
  np.testing.assert_allclose(mu, g['rand_mu'], rtol=0, atol=1e-10)
  np.testing.assert_allclose(sd ** 2, g['rand_sd'] ** 2, rtol=0, atol=1e-8)
  np.random.seed(5)
  tag, cts, dscr, other, probs = hp_grid.fit_gp(X, Y, layout, g['rand_exp_sampling_bounds'],
                                                g['rand_exp_sampling_dscr_vals'], method='rand_exp_sampling',
                                                max_evals=int(g['rand_exp_sampling_max_evals']))
  assert tag == 'sample_hps_with_probs' and len(other) == len(cts)
  assert (cts == g['rand_exp_sampling_cts']).all() and (np.array(dscr) == g['rand_exp_sampling_dscr']).all()
  np.testing.assert_allclose(probs, g['rand_exp_sampling_probs'], rtol=1e-7, atol=1e-12)


def test_pdoo_over_the_hyperparameters(setup):
  hp_grid, g, layout = setup
  X, Y = g['X'], g['Y']
  tag, gp, (cts, dscr) = hp_grid.fit_gp(X, Y, layout, g['pdoo_bounds'], g['pdoo_dscr_vals'], method='pdoo',
                                        max_evals=int(g['pdoo_max_evals']))
  assert tag == 'fitted_gp' and len(cts) == 6 and dscr[0] in (0.5, 1.5, 2.5)
  same = (np.array(cts) == g['pdoo_cts']).all() and (np.array(dscr) == g['pdoo_dscr']).all()
  # a tree search may branch differently on a 1e-11 LML difference; then it must still end at least as high
  assert same or gp.compute_log_marginal_likelihood() >= float(g['pdoo_lml']) - 1e-6
  if same:
    np.testing.assert_allclose(gp.compute_log_marginal_likelihood(), float(g['pdoo_lml']), rtol=1e-10)
