"""
Seeded synthetic workloads for the parity tests and bench.py (BASELINE.json configs).

Vectorised restatements of the standard benchmark objectives Dragonfly ships as data generators
(dragonfly/utils/euclidean_synthetic_functions.py: Hartmann-6 :37-49/:16-19, Branin :108-149,
Borehole :152-192, Park1 :195-211, '<name>-<dim>' tiling :254-284).  Inputs are in the unit cube
(EuclideanFunctionCaller normalises, experiment_caller.py:399-400); Y is never standardised.
Pure NumPy, host side, not on the timed path.
"""
import numpy as np

_H6_A = np.array([[10, 3, 17, 3.5, 1.7, 8],
                  [0.05, 10, 17, 0.1, 8, 14],
                  [3, 3.5, 1.7, 10, 17, 8],
                  [17, 8, 0.05, 10, 0.1, 14]], dtype=np.float64)
_H6_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886],
                         [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650],
                         [4047, 8828, 8732, 5743, 1091, 381]], dtype=np.float64)
_H6_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_H6_MAX = 3.322368

_BOREHOLE_BOUNDS = np.array([[0.05, 0.15], [100, 50000], [63070, 115600], [990, 1110],
                             [63.1, 116], [700, 820], [1120, 1680], [9855, 12045]], dtype=np.float64)
_BRANIN_BOUNDS = np.array([[-5, 10], [0, 15]], dtype=np.float64)


def hartmann6(X):
  """ X: (n, 6) in [0,1]^6 -> (n,) ; maximisation form, capped at the known optimum. """
  X = np.asarray(X, dtype=np.float64)
  inner = (_H6_A[None, :, :] * (_H6_P[None, :, :] - X[:, None, :]) ** 2).sum(axis=2)
  return np.minimum(_H6_MAX, np.exp(-inner).dot(_H6_ALPHA))


def branin(X_unit):
  """ X_unit: (n, 2) in the unit square, mapped to [-5,10]x[0,15]; returns -branin (maximise). """
  X = np.asarray(X_unit, dtype=np.float64) * (_BRANIN_BOUNDS[:, 1] - _BRANIN_BOUNDS[:, 0]) \
      + _BRANIN_BOUNDS[:, 0]
  a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6.0, 10.0, 1 / (8 * np.pi)
  x1, x2 = X[:, 0], X[:, 1]
  return -(a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s)


def borehole(Z_unit, X_unit):
  """ Multi-fidelity Borehole: Z_unit (n,1) fidelity in [0,1], X_unit (n,8) in the unit cube. """
  X = np.asarray(X_unit, dtype=np.float64) * (_BOREHOLE_BOUNDS[:, 1] - _BOREHOLE_BOUNDS[:, 0]) \
      + _BOREHOLE_BOUNDS[:, 0]
  z = np.asarray(Z_unit, dtype=np.float64).reshape(-1)
  rw, r, Tu, Hu, Tl, Hl, L, Kw = [X[:, i] for i in range(8)]
  lg = np.log(r / rw)
  frac2 = 2 * L * Tu / (lg * rw ** 2 * Kw)
  f2 = np.minimum(309.523221, 2 * np.pi * Tu * (Hu - Hl) / (lg * (1 + frac2 + Tu / Tl)))
  f1 = 5 * Tu * (Hu - Hl) / (lg * (1.5 + frac2 + Tu / Tl))
  return f2 * z + f1 * (1 - z)


def park1(X):
  """ X: (n, 4) in [0,1]^4. """
  X = np.asarray(X, dtype=np.float64)
  x1, x2, x3, x4 = [X[:, i] for i in range(4)]
  ret1 = (x1 / 2) * (np.sqrt(1 + (x2 + x3 ** 2) * x4 / (x1 ** 2)) - 1)
  ret2 = (x1 + 3 * x4) * np.exp(1 + np.sin(x3))
  return np.minimum(ret1 + ret2, 25.5872304)


def tiled(func, group_dim, X):
  """ '<name>-<dim>': sum of func over consecutive full groups of group_dim coordinates
      (the trailing remainder coordinates do not enter the objective). """
  X = np.asarray(X, dtype=np.float64)
  num_groups = X.shape[1] // group_dim
  out = np.zeros(len(X))
  for j in range(num_groups):
    out += func(X[:, j * group_dim:(j + 1) * group_dim])
  return out


def make_workload(name, n_train=None, n_cand=None, seed_train=0, seed_cand=1):
  """ The five BASELINE.json configs (+ the headline N=5000 case) with the fixed hyper-parameters
      of SURVEY.md 8d.  Returns a dict of plain NumPy arrays / floats; kernels are described as
      nested dicts understood by dragonfly_b200.kernel.kernel_from_spec and by the tests' oracle
      builder. """
  rs = np.random.RandomState(seed_train)
  rc = np.random.RandomState(seed_cand)
  if name == 'c1_branin_se_ei':
    n, m, d = n_train or 50, n_cand or 10000, 2
    X = rs.random_sample((n, d)); Y = branin(X)
    kern = dict(type='se', dim=d, scale=float(Y.var()), dim_bandwidths=[0.2] * d)
    acq = dict(name='ei', curr_best=float(Y.max()))
    mean = float(np.median(Y))
  elif name in ('c2_hartmann6_matern_ucb', 'headline_hartmann6_matern_ei'):
    head = name.startswith('headline')
    n, m, d = n_train or (5000 if head else 2000), n_cand or 1000000, 6
    X = rs.random_sample((n, d)); Y = hartmann6(X)
    kern = dict(type='matern', dim=d, nu=2.5, scale=float(Y.var()), dim_bandwidths=[0.3] * d)
    acq = dict(name='ei', curr_best=float(Y.max())) if head else dict(name='ucb', t=n)
    mean = float(np.median(Y))
  elif name == 'c3_additive40_add_ucb':
    n, m, d = n_train or 5000, n_cand or 4000000, 40
    X = rs.random_sample((n, d)); Y = tiled(hartmann6, 6, X)
    groups = [list(range(6 * j, min(6 * j + 6, d))) for j in range(7)]
    kern = dict(type='additive', scale=float(Y.var()) / 7.0, groupings=groups,
                kernels=[dict(type='matern', dim=len(g), nu=2.5, scale=1.0,
                              dim_bandwidths=[0.5] * len(g)) for g in groups])
    acq = dict(name='add_ucb', t=n)
    mean = float(np.median(Y))
  elif name == 'c4_borehole_mf_ucb':
    n, m, d = n_train or 4000, n_cand or 1000000, 8
    Z = rs.random_sample((n, 1)); Xd = rs.random_sample((n, d)); Y = borehole(Z, Xd)
    X = np.concatenate((Z, Xd), axis=1)
    kern = dict(type='coordinate_product', dim=1 + d, scale=float(Y.var()),
                coordinate_list=[[0], list(range(1, 1 + d))],
                kernels=[dict(type='se', dim=1, scale=1.0, dim_bandwidths=[0.7]),
                         dict(type='se', dim=d, scale=1.0, dim_bandwidths=[0.4] * d)])
    acq = dict(name='ucb', t=n, fidel_to_opt=[1.0])
    mean = float(np.median(Y))
  elif name == 'c5_park1_20_ts':
    n, m, d = n_train or 5000, n_cand or 1000000, 20
    X = rs.random_sample((n, d)); Y = tiled(park1, 4, X)
    kern = dict(type='matern', dim=d, nu=2.5, scale=float(Y.var()), dim_bandwidths=[0.5] * d)
    acq = dict(name='ts', num_draws=256)
    mean = float(np.median(Y))
  else:
    raise ValueError('unknown workload %s' % (name))
  noise_var = 0.01 * float(Y.var())
  cand_dim = d
  cands = rc.random_sample((m, cand_dim))
  return dict(name=name, X=X, Y=Y, kernel=kern, acq=acq, mean_const=mean, noise_var=noise_var,
              candidates=cands, dim=d, n_train=n, n_cand=m)
