"""
The hyper-parameter "grid" objective of GPFitter on the device:

  GPFitter._tuning_objective (gp_core.py:551-563)  = LML of the GP built from one hp vector
  GPFitter.build_gp (gp_core.py:501-543)           : [mean const if mean_func_type == 'tune'],
                                                     [log noise if noise_var_type == 'tune'], child hps
  EuclideanGPFitter._child_build_gp (euclidean_gp.py:325-339) /
  get_euclidean_integral_gp_kernel_with_scale (:808-900): log scale, log bandwidth (x d or x 1),
                                                     discrete nu for Matern
  _rand_exp_sampling_wrap (gp_core.py:439-445)     : probs = exp(lml - max lml), normalised

Each objective evaluation is one DFB_BUILD_LML_ONLY factorisation (K build + blocked Cholesky with
the y row riding along; no L^-1, no alpha).  The samples are independent, so under torch.distributed
they shard across ranks with one all-gather of the LML values at the end (NCCL on GPUs, gloo in the
CPU tests of the sharding logic).
"""
import numpy as np

from . import _lib
from .kernel import SEKernel, MaternKernel, AdditiveKernel, CoordinateProductKernel, build_descriptor
from .gp_core import stable_cholesky_on_device


class EuclideanHPLayout(object):
  """ How a continuous hp vector maps to (mean const, noise var, kernel) for an SE / Matern GP. """

  def __init__(self, dim, kernel_type='matern', nu=2.5, use_same_bandwidth=False,
               mean_func_type='median', mean_func_const=0.0, noise_var_type='tune',
               noise_var_label=0.05, noise_var_value=0.1, use_additive_gp=False, add_max_group_size=6,
               num_groups_per_group_size=-1):
    if kernel_type not in ('se', 'matern'):
      raise NotImplementedError('kernel_type %s is outside the B200 hot-path scope.' % (kernel_type))
    self.dim, self.kernel_type, self.nu = dim, kernel_type, nu
    self.use_same_bandwidth = use_same_bandwidth
    self.mean_func_type, self.mean_func_const = mean_func_type, mean_func_const
    self.noise_var_type = noise_var_type
    self.noise_var_label, self.noise_var_value = noise_var_label, noise_var_value
    # additive models (euclidean_gp.py:50-60, 243-248): the group size is one more discrete hyper-parameter and
    # every objective evaluation carries a random grouping of the coordinates
    self.use_additive_gp = use_additive_gp
    self.add_max_group_size = min(add_max_group_size, dim)
    self.num_groups_per_group_size = num_groups_per_group_size

  def num_hps(self):
    n = 1 + (1 if self.use_same_bandwidth else self.dim)
    n += 1 if self.mean_func_type == 'tune' else 0
    n += 1 if self.noise_var_type == 'tune' else 0
    return n

  def bounds(self, X, Y, tune_nu=None):
    """ (cts_hp_bounds, dscr_hp_vals) the way the reference's fitter sets them up from the data
        (gp_core.py:336-338, 396-416; euclidean_gp.py:253-276): mean value (if tuned), log noise (if tuned),
        log scale, log bandwidth(s); discrete [0.5, 1.5, 2.5] for a Matern kernel whose nu is tuned
        (options.matern_nu < 0; default here: tuned iff self.nu is None or negative). """
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    Y_var = Y.std() ** 2 + 0.0001 if len(Y) > 0 else 0.0001
    out = []
    if self.mean_func_type == 'tune':
      Y_std = np.sqrt(Y_var)
      Y_median = np.median(Y) if len(Y) > 0 else 0.0
      Y_half_range = 0.5 * (max(Y) - min(Y)) if len(Y) > 0 else 1.0
      Y_width = 0.5 * (Y_half_range + Y_std)
      out.append([Y_median - 3 * Y_width, Y_median + 3 * Y_width])
    if self.noise_var_type == 'tune':
      out.append([np.log(0.005 * Y_var), np.log(0.2 * Y_var)])
    out.append([np.log(0.1 * Y_var), np.log(10 * Y_var)])
    X_std_norm = np.linalg.norm(X, 'fro') + 1e-4
    single = [np.log(0.01 * X_std_norm), np.log(10 * X_std_norm)]
    out += [single] * (1 if self.use_same_bandwidth else self.dim)
    if tune_nu is None:
      tune_nu = self.kernel_type == 'matern' and (self.nu is None or self.nu < 0)
    dscr = [[0.5, 1.5, 2.5]] if (self.kernel_type == 'matern' and tune_nu) else []
    if self.use_additive_gp:
      dscr.append([x + 1 for x in range(self.add_max_group_size)])
    return np.array(out), dscr

  def _mean_and_noise(self, hp, Y):
    """ Pops the mean value / log noise from the front of `hp` when they are tuned (gp_core.py:509-538). """
    if self.mean_func_type == 'mean':
      mean_const = np.mean(Y)
    elif self.mean_func_type == 'median':
      mean_const = np.median(Y)
    elif self.mean_func_type == 'upper_bound':
      mean_const = np.mean(Y) + 3 * np.std(Y)
    elif self.mean_func_type == 'const':
      mean_const = self.mean_func_const
    elif self.mean_func_type == 'tune':
      mean_const = hp.pop(0)
    else:
      mean_const = 0
    if self.noise_var_type == 'tune':
      noise_var = np.exp(hp.pop(0))
    elif self.noise_var_type == 'label':
      noise_var = self.noise_var_label * (Y.std() ** 2)
    else:
      noise_var = self.noise_var_value
    return mean_const, noise_var

  def unpack(self, hp, Y, nu=None, groupings=None):
    """ gp_core.py:509-538 + euclidean_gp.py:801-861 """
    hp = list(np.asarray(hp, dtype=np.float64))
    Y = np.asarray(Y, dtype=np.float64)
    mean_const, noise_var = self._mean_and_noise(hp, Y)
    scale = np.exp(hp.pop(0))
    if self.use_same_bandwidth:
      bws = [np.exp(hp.pop(0))] * self.dim
    else:
      bws = [np.exp(hp.pop(0)) for _ in range(self.dim)]
    assert len(hp) == 0
    nu = self.nu if nu is None else nu
    if groupings is not None:
      # get_euclidean_integral_gp_kernel_with_scale (euclidean_gp.py:826-831, 850-861, 895-897): group kernels
      # with scale 1 on their own bandwidths, the outer scale on the sum
      groups = [[int(i) for i in grp] for grp in groupings]
      make = (lambda grp: SEKernel(len(grp), 1.0, [bws[i] for i in grp])) if self.kernel_type == 'se' else \
             (lambda grp: MaternKernel(len(grp), nu, 1.0, [bws[i] for i in grp]))
      kern = AdditiveKernel(scale, [make(grp) for grp in groups], groups)
    elif self.kernel_type == 'se':
      kern = SEKernel(self.dim, scale, bws)
    else:
      kern = MaternKernel(self.dim, nu, scale, bws)
    return float(mean_const), float(noise_var), kern


class EuclideanMFHPLayout(EuclideanHPLayout):
  """ Hyper-parameter vector of the reference's EuclideanMFGPFitter (euclidean_gp.py:432-483, 680-709): [mean const]?
      [log noise]? log scale, log fidelity bandwidth(s), log domain bandwidth(s); at most one tuned Matern nu (fidelity
      or domain) as the discrete hp.  The GP lives on [z || x] rows with the product kernel
      scale * k_F(z, z') * k_D(x, x') (fidelity and domain kernels with scale 1). """

  def __init__(self, fidel_dim, domain_dim, fidel_kernel_type='se', domain_kernel_type='se', fidel_nu=2.5,
               domain_nu=2.5, fidel_use_same_bandwidth=False, domain_use_same_bandwidth=False, **kwargs):
    for kt in (fidel_kernel_type, domain_kernel_type):
      if kt not in ('se', 'matern'):
        raise NotImplementedError('kernel_type %s is outside the B200 hot-path scope.' % (kt))
    super(EuclideanMFHPLayout, self).__init__(fidel_dim + domain_dim, 'se', **kwargs)
    self.fidel_dim, self.domain_dim = fidel_dim, domain_dim
    self.fidel_kernel_type, self.domain_kernel_type = fidel_kernel_type, domain_kernel_type
    self.fidel_nu, self.domain_nu = fidel_nu, domain_nu
    self.fidel_use_same_bandwidth = fidel_use_same_bandwidth
    self.domain_use_same_bandwidth = domain_use_same_bandwidth
    if self.use_additive_gp:
      raise NotImplementedError('Additive domain kernels in the MF fitter are outside the device path.')

  def tuned_nus(self):
    return [self.fidel_kernel_type == 'matern' and self.fidel_nu < 0,
            self.domain_kernel_type == 'matern' and self.domain_nu < 0]

  def num_hps(self):
    n = 1 + (1 if self.fidel_use_same_bandwidth else self.fidel_dim)
    n += 1 if self.domain_use_same_bandwidth else self.domain_dim
    n += 1 if self.mean_func_type == 'tune' else 0
    n += 1 if self.noise_var_type == 'tune' else 0
    return n

  def bounds(self, X, Y, tune_nu=None):
    raise NotImplementedError('Use the bounds of the reference EuclideanMFGPFitter (fitter.cts_hp_bounds).')

  def unpack(self, hp, Y, nu=None, groupings=None):
    """ gp_core.py:509-538 + euclidean_gp.py:680-709 """
    hp = list(np.asarray(hp, dtype=np.float64))
    Y = np.asarray(Y, dtype=np.float64)
    mean_const, noise_var = self._mean_and_noise(hp, Y)
    scale = np.exp(hp.pop(0))

    def bandwidths(n, same):
      return [np.exp(hp.pop(0))] * n if same else [np.exp(hp.pop(0)) for _ in range(n)]
    f_bws = bandwidths(self.fidel_dim, self.fidel_use_same_bandwidth)
    d_bws = bandwidths(self.domain_dim, self.domain_use_same_bandwidth)
    assert len(hp) == 0
    f_tuned, d_tuned = self.tuned_nus()
    assert not (f_tuned and d_tuned), 'at most one tuned Matern nu on the device path'
    f_nu = nu if f_tuned else self.fidel_nu
    d_nu = nu if d_tuned else self.domain_nu
    k_f = SEKernel(self.fidel_dim, 1.0, f_bws) if self.fidel_kernel_type == 'se' else \
        MaternKernel(self.fidel_dim, f_nu, 1.0, f_bws)
    k_d = SEKernel(self.domain_dim, 1.0, d_bws) if self.domain_kernel_type == 'se' else \
        MaternKernel(self.domain_dim, d_nu, 1.0, d_bws)
    fidel_coords = list(range(self.fidel_dim))
    domain_coords = list(range(self.fidel_dim, self.fidel_dim + self.domain_dim))
    kern = CoordinateProductKernel(self.dim, scale, [k_f, k_d], [fidel_coords, domain_coords])
    return float(mean_const), float(noise_var), kern


# One LML-only build at N = 5000 is a 40-step dependency chain (chol_diag -> panel -> next column) that leaves most
# of the GPU idle (7.7 ms for 42 GFLOP); the hp samples are independent, so `lanes` of them are built
# concurrently -- one DevicePosterior (handle + workspace) per lane, each on its own CUDA stream, driven from its
# own host thread (ctypes releases the GIL during the C call).  Sample i goes to lane i % lanes; every LML is
# computed by the same kernels whatever the lane count, so the values do not depend on it.
DEFAULT_LANES = 3


def _lane_worker(lane_post, stream, X, Y, hps, idxs, layout, nus, out, groupings=None):
  import torch
  with torch.cuda.device(lane_post.device), torch.cuda.stream(stream):
    lane_post.bind_current_stream()
    last_mean = None
    for i in idxs:
      mean_const, noise_var, kern = layout.unpack(hps[i], Y, None if nus is None else nus[i],
                                                  None if groupings is None else groupings[i])
      if last_mean is None or mean_const != last_mean:
        lane_post.set_train(X, Y - mean_const)
        last_mean = mean_const
      lane_post.set_kernel(build_descriptor(kern, train_dim=X.shape[1], cand_dim=X.shape[1]))
      out[i], _ = stable_cholesky_on_device(lane_post, noise_var, flags=_lib.DFB_BUILD_LML_ONLY)


def lml_for_hyperparams(X, Y, hps, layout, nus=None, post=None, device=None, lanes=None, groupings=None):
  """ LML of the GP built from each hp vector (rows of `hps`); `nus` optionally gives the discrete
      Matern nu per sample.  Returns (lmls, post) -- `post` can be passed back in to reuse the
      device workspaces (it carries the extra lanes). """
  import threading
  import torch
  from .device import DevicePosterior
  X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
  Y = np.asarray(Y, dtype=np.float64)
  if post is None or post.n_max < len(X):
    post = DevicePosterior(len(X), device=device)
  lmls = np.empty(len(hps))
  if lanes is None:
    lanes = DEFAULT_LANES if len(hps) >= 2 * DEFAULT_LANES else 1
  lanes = max(1, min(int(lanes), len(hps)))
  if lanes == 1:
    _lane_worker(post, torch.cuda.current_stream(post.device), X, Y, hps, range(len(hps)), layout, nus, lmls,
                 groupings)
    return lmls, post
  extra = getattr(post, '_hp_lanes', [])
  while len(extra) < lanes - 1:
    extra.append((DevicePosterior(len(X), device=post.device.index), torch.cuda.Stream(post.device)))
  post._hp_lanes = extra
  main_stream = torch.cuda.current_stream(post.device)
  workers, errors = [], []

  def guarded(*a):
    try:
      _lane_worker(*a)
    except BaseException as e:  # pylint: disable=broad-except
      errors.append(e)
  for j in range(1, lanes):
    lane_post, stream = extra[j - 1]
    stream.wait_stream(main_stream)
    t = threading.Thread(target=guarded, args=(lane_post, stream, X, Y, hps, range(j, len(hps), lanes), layout,
                                               nus, lmls, groupings))
    t.start()
    workers.append(t)
  guarded(post, main_stream, X, Y, hps, range(0, len(hps), lanes), layout, nus, lmls, groupings)
  for t in workers:
    t.join()
  for j in range(1, lanes):
    main_stream.wait_stream(extra[j - 1][1])
  if errors:
    raise errors[0]
  return lmls, post


def default_max_evals(method, num_hps):
  """ gp_core.py:456-462 """
  if method in ('direct', 'pdoo'):
    return int(min(1e4, max(500, num_hps * 50)))
  if method == 'rand':
    return int(min(1e4, max(500, num_hps * 200)))
  if method == 'rand_exp_sampling':
    return int(min(1e5, max(500, num_hps * 400)))
  raise ValueError('Unknown ml_hp_tune_opt method %s.' % (method))


def fit_gp(X, Y, layout, cts_hp_bounds, dscr_hp_vals=(), method='rand_exp_sampling', max_evals=None,
           device=None, gp_factory=None, build_gp=None):
  """ GPFitter.fit_gp for hp_tune_criterion == 'ml' (gp_core.py:783-808) with the marginal likelihood of every
      hyper-parameter vector evaluated on the device, in batches instead of one _tuning_objective call at a time
      (gp_core.py:551-563):
        'rand'               random_maximise over the continuous hps for every combination of the discrete ones
                             (oper_utils.py:69-80, gp_core.py:787-799): all max_evals candidates of a combination are
                             one lml_for_hyperparams call (concurrent build lanes);
        'pdoo' / 'direct'    pdoo_maximise (oper_utils.py:257-271; 'direct' is PDOO wherever the reference's Fortran
                             DIRECT is not built) through dragonfly_b200.doo: both children of a split in one call;
        'rand_exp_sampling'  random_sample_cts_dscr + exp(lml - max) weights (oper_utils.py:362-371,
                             gp_core.py:439-445): one call for all samples.
      The global NumPy RNG is consumed exactly like the reference does, so a seeded run picks the same
      hyper-parameters.  `cts_hp_bounds` / `dscr_hp_vals` are the fitter's (euclidean_gp.py:222-320); for this layout
      the only discrete hp is the Matern nu.  Returns what the reference returns:
        ('fitted_gp', gp, (cts_hps, dscr_hps))   or   ('sample_hps_with_probs', cts, dscr, [None] * n, probs). """
  from itertools import product as itertools_product
  from .gp_core import GP, ConstantMean
  from .gpb_acquisitions import map_to_bounds, _reference_fortran_direct_available
  X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
  Y = np.asarray(Y, dtype=np.float64)
  bounds = np.asarray(cts_hp_bounds, dtype=np.float64)
  dscr_hp_vals = [list(v) for v in dscr_hp_vals]
  additive = bool(getattr(layout, 'use_additive_gp', False))
  if len(dscr_hp_vals) > (2 if additive else 1):
    raise NotImplementedError('Discrete hyper-parameters on the device path: the Matern nu and, for additive '
                              'models, the group size.')
  has_nu = len(dscr_hp_vals) == (2 if additive else 1)       # [nu]? then [group size]? (euclidean_gp.py:226-248)
  dim = layout.dim
  if max_evals is None:
    max_evals = default_max_evals(method, len(bounds) + len(dscr_hp_vals))
  n_evals = int(max_evals)
  state = {'post': None}

  def lmls_of(hps, nus, groupings=None):
    vals, state['post'] = lml_for_hyperparams(X, Y, hps, layout, nus=nus, post=state['post'], device=device,
                                              groupings=groupings)
    return vals

  def build(cts, dscr, groupings=None):
    if build_gp is not None:                         # e.g. the reference fitter's own build_gp (fit_gp_on_fitter)
      return build_gp(cts, dscr, groupings)
    mean_const, noise_var, kern = layout.unpack(cts, Y, dscr[0] if has_nu else None, groupings)
    make = GP if gp_factory is None else gp_factory
    return make(list(X), list(Y), kern, ConstantMean(mean_const), noise_var)

  def random_grouping(group_size):
    rand_perm = list(np.random.permutation(dim))               # euclidean_gp.py:733-735, 760-761
    return [rand_perm[i:i + group_size] for i in range(0, dim, group_size)]

  if method == 'rand_exp_sampling':
    if additive:
      # sample_cts_dscr_hps_for_rand_exp_sampling_in_add_model (euclidean_gp.py:749-776): per sample a group size, a
      # random grouping, the discrete hps (group size overwritten), the continuous hps -- in that RNG order; the
      # weights are exp(lml) / sum WITHOUT subtracting the maximum, as written there
      cts, dscr, groupings = [], [], []
      for _ in range(n_evals):
        group_size = np.random.choice(dscr_hp_vals[-1])
        groupings.append(random_grouping(group_size))
        cur = [np.random.choice(categ) for categ in dscr_hp_vals]
        cur[-1] = group_size
        dscr.append(cur)
        cts.append(map_to_bounds(np.random.random((len(bounds),)), bounds))
      vals = lmls_of(np.array(cts), [d[0] for d in dscr] if has_nu else None, groupings)
      probs = np.exp(vals)
      from argparse import Namespace
      other = [Namespace(add_gp_groupings=grp) for grp in groupings]      # as the reference returns them (:762)
      return 'sample_hps_with_probs', cts, dscr, other, probs / probs.sum()
    cts = map_to_bounds(np.random.random((n_evals, len(bounds))), bounds)
    dscr = [[np.random.choice(categ) for categ in dscr_hp_vals] for _ in range(n_evals)]
    vals = lmls_of(cts, [d[0] for d in dscr] if dscr_hp_vals else None)
    return 'sample_hps_with_probs', cts, dscr, [None] * n_evals, rand_exp_sampling_probs(vals)
  if method == 'direct' and _reference_fortran_direct_available():
    raise NotImplementedError('Fortran DIRECT is a sequential host optimiser; use pdoo / rand / rand_exp_sampling.')
  if method not in ('rand', 'pdoo', 'direct'):
    raise ValueError('Unknown ml_hp_tune_opt method %s.' % (method))

  def optimise_cts(nu, evals, groupings=None):
    """ cts_hp_optimise (gp_core.py:463-472) for one setting of the discrete hps (and one grouping). """
    rep = (lambda k: None) if groupings is None else (lambda k: [groupings] * k)
    if method == 'rand':
      pts = map_to_bounds(np.random.random((int(evals), len(bounds))), bounds)
      vals = lmls_of(pts, None if nu is None else [nu] * len(pts), rep(len(pts)))
      idx = int(np.argmax(vals))
      return vals[idx], pts[idx]
    from .doo import pdoo_maximise
    val, pt, _ = pdoo_maximise(lambda P: lmls_of(P, None if nu is None else [nu] * len(P), rep(len(P))), bounds,
                               evals)
    return val, pt

  best_val, best_cts, best_dscr, best_groupings = -np.inf, None, None, None
  for dscr in itertools_product(*dscr_hp_vals):
    nu = dscr[0] if has_nu else None
    if not additive:
      opt_val, opt_pt, opt_groupings = optimise_cts(nu, max_evals) + (None,)
    else:
      # optimise_cts_hps_for_given_dscr_hps_in_add_model (euclidean_gp.py:718-746)
      group_size = dscr[-1]
      n_groupings = layout.num_groups_per_group_size
      if n_groupings < 0:
        n_groupings = 1 if group_size == 1 else max(5, min(2 * dim, 25))
      opt_val, opt_pt, opt_groupings = -np.inf, None, None
      for _ in range(n_groupings):
        groupings = random_grouping(group_size)
        val, pt = optimise_cts(nu, int(max(500, max_evals / n_groupings)), groupings)
        if val > opt_val:
          opt_val, opt_pt, opt_groupings = val, pt, groupings
    if opt_val > best_val:
      best_val, best_cts, best_dscr, best_groupings = opt_val, list(opt_pt), list(dscr), opt_groupings
  return 'fitted_gp', build(best_cts, best_dscr, best_groupings), (best_cts, best_dscr)


# ---- drop-in for GPFitter.fit_gp on a Dragonfly EuclideanGPFitter --------------------------------
def layout_from_fitter(fitter):
  """ The EuclideanHPLayout equivalent to a reference EuclideanGPFitter's options (euclidean_gp.py:205-248,
      gp_core.py:509-538), or None when the fitter tunes something the device path does not cover. """
  opt = fitter.options
  if getattr(opt, 'mean_func', None) is not None:
    return None
  common = dict(mean_func_type=opt.mean_func_type, mean_func_const=getattr(opt, 'mean_func_const', 0.0),
                noise_var_type=opt.noise_var_type, noise_var_label=getattr(opt, 'noise_var_label', 0.05),
                noise_var_value=getattr(opt, 'noise_var_value', 0.1))
  if hasattr(fitter, 'fidel_dim') and hasattr(opt, 'fidel_kernel_type'):
    # EuclideanMFGPFitter (euclidean_gp.py:418-716)
    if (opt.fidel_kernel_type not in ('se', 'matern') or opt.domain_kernel_type not in ('se', 'matern') or
        getattr(opt, 'domain_use_additive_gp', False)):
      return None
    layout = EuclideanMFHPLayout(
        fitter.fidel_dim, fitter.domain_dim, opt.fidel_kernel_type, opt.domain_kernel_type,
        fidel_nu=getattr(opt, 'fidel_matern_nu', 2.5), domain_nu=getattr(opt, 'domain_matern_nu', 2.5),
        fidel_use_same_bandwidth=bool(getattr(opt, 'fidel_use_same_bandwidth', False)),
        domain_use_same_bandwidth=bool(getattr(opt, 'domain_use_same_bandwidth', False)), **common)
    return None if sum(layout.tuned_nus()) > 1 else layout
  kernel_type = getattr(fitter, 'kernel_type', getattr(opt, 'kernel_type', None))
  if kernel_type not in ('se', 'matern'):
    return None
  additive = bool(getattr(opt, 'use_additive_gp', False))
  return EuclideanHPLayout(
      fitter.dim, kernel_type, nu=getattr(opt, 'matern_nu', 2.5),
      use_same_bandwidth=bool(getattr(opt, 'use_same_bandwidth', False)),
      mean_func_type=opt.mean_func_type, mean_func_const=getattr(opt, 'mean_func_const', 0.0),
      noise_var_type=opt.noise_var_type, noise_var_label=getattr(opt, 'noise_var_label', 0.05),
      noise_var_value=getattr(opt, 'noise_var_value', 0.1), use_additive_gp=additive,
      add_max_group_size=getattr(fitter, 'add_max_group_size', getattr(opt, 'add_max_group_size', 6)),
      num_groups_per_group_size=getattr(opt, 'num_groups_per_group_size', -1))


def fit_gp_on_fitter(fitter, reference_fit_gp, num_samples=1, hp_tune_criterion=None):
  """ GPFitter.fit_gp (gp_core.py:783-821) for a reference EuclideanGPFitter instance: hp_tune_criterion 'ml' with
      ml_hp_tune_opt rand / rand_exp_sampling / pdoo / direct-without-Fortran runs fit_gp above (every batch of
      _tuning_objective evaluations as one lml_for_hyperparams call; same global-RNG consumption, same selection,
      the final GP built by the fitter's own build_gp); everything else is handed to `reference_fit_gp`. """
  from argparse import Namespace
  from .gpb_acquisitions import _reference_fortran_direct_available
  crit = fitter.options.hp_tune_criterion if hp_tune_criterion is None else hp_tune_criterion
  method = getattr(fitter, 'ml_hp_tune_opt_method', None)
  layout = layout_from_fitter(fitter) if crit == 'ml' else None
  if (layout is None or method not in ('rand', 'rand_exp_sampling', 'pdoo', 'direct') or
      (method == 'direct' and _reference_fortran_direct_available())):
    return reference_fit_gp(fitter, num_samples, hp_tune_criterion)
  other = lambda grp: None if grp is None else Namespace(add_gp_groupings=grp)
  if isinstance(layout, EuclideanMFHPLayout):
    X_mat = np.concatenate((np.asarray(fitter.ZZ, dtype=np.float64).reshape(len(fitter.YY), -1),
                            np.asarray(fitter.XX, dtype=np.float64).reshape(len(fitter.YY), -1)), axis=1)
    Y_vec = np.array(fitter.YY)
  else:
    X_mat, Y_vec = np.array(fitter.X), np.array(fitter.Y)
  return fit_gp(X_mat, Y_vec, layout, fitter.cts_hp_bounds, fitter.dscr_hp_vals,
                method=method, max_evals=fitter.hp_tune_max_evals,
                build_gp=lambda cts, dscr, grp: fitter.build_gp(cts, dscr, other_gp_params=other(grp)))


def bind_fit_gp(fitter_class):
  """ Re-binds fit_gp on a reference GPFitter class (dragonfly.gp.gp_core.GPFitter); returns the original. """
  original = fitter_class.fit_gp

  def fit_gp_b200(self, num_samples=1, hp_tune_criterion=None):
    return fit_gp_on_fitter(self, original, num_samples, hp_tune_criterion)
  fitter_class.fit_gp = fit_gp_b200
  return original


def rand_exp_sampling_probs(lml_vals):
  """ gp_core.py:443-444 """
  lml_vals = np.asarray(lml_vals, dtype=np.float64)
  probs = np.exp(lml_vals - max(lml_vals))
  return probs / probs.sum()


def sharded_lml_grid(X, Y, hps, layout, nus=None, device=None, group=None):
  """ Rank r evaluates hps[lo:hi]; one all-gather of the fp64 LML values.  Every rank returns the
      full vector and the rand_exp_sampling probabilities. """
  import torch
  import torch.distributed as dist
  from .dist import shard_bounds
  hps = np.asarray(hps, dtype=np.float64)
  H = len(hps)
  if not (dist.is_available() and dist.is_initialized()):
    lmls, _ = lml_for_hyperparams(X, Y, hps, layout, nus, device=device)
    return lmls, rand_exp_sampling_probs(lmls)
  rank, world = dist.get_rank(group), dist.get_world_size(group)
  lo, hi = shard_bounds(H, rank, world)
  mine, _ = lml_for_hyperparams(X, Y, hps[lo:hi], layout, None if nus is None else nus[lo:hi],
                                device=device) if hi > lo else (np.empty(0), None)
  return gather_shards(mine, H, group=group, device=device)


def gather_shards(mine, total, group=None, device=None):
  """ All-gather of variable-length fp64 shards (padded to the largest shard). """
  import torch
  import torch.distributed as dist
  from .dist import shard_bounds
  world = dist.get_world_size(group)
  backend = dist.get_backend(group)
  dev = torch.device('cpu') if backend == 'gloo' else (
      torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device('cuda', device))
  cap = max(shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world))
  buf = torch.zeros(max(cap, 1), dtype=torch.float64, device=dev)
  buf[:len(mine)] = torch.from_numpy(np.asarray(mine, dtype=np.float64)).to(dev)
  out = [torch.empty_like(buf) for _ in range(world)]
  dist.all_gather(out, buf, group=group)
  parts = []
  for r in range(world):
    lo, hi = shard_bounds(total, r, world)
    parts.append(out[r][:hi - lo].cpu().numpy())
  lmls = np.concatenate(parts) if parts else np.empty(0)
  return lmls, rand_exp_sampling_probs(lmls)
