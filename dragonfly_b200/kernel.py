"""
Host-side mirror of dragonfly/gp/kernel.py for the kernels on the hot path: SEKernel, MaternKernel,
AdditiveKernel, CoordinateProductKernel (same constructor arguments, `hyperparams` dict, `dim`,
`kernel_list` / `groupings` / `coordinate_list`, `is_guaranteed_psd`, `__call__`), plus the
translation of such kernel objects -- these classes OR the reference's own, duck-typed by class name
so that a patched Dragonfly install needs no changes -- into the POD `dfb_kernel_desc` that
libdfb200's CUDA kernels evaluate (include/dfb200.h).

`Kernel.__call__(X1, X2)` (kernel.py:72-83) runs on the GPU through dfb_kernel_matrix.  Kernel types
outside the hot-path scope (Poly, ExpDecay, Hamming, ESP, NN kernels) raise NotImplementedError:
there is no CPU fallback.
"""
import math

import numpy as np

from . import _lib


# ---------------------------------------------------------------------------------------------
# The Kernel surface (kernel.py:59-129)
# ---------------------------------------------------------------------------------------------
class Kernel(object):
  """ kernel.py:59-129 """

  def __init__(self):
    super(Kernel, self).__init__()
    self.hyperparams = {}

  def is_guaranteed_psd(self):
    raise NotImplementedError('Implement in a child class.')

  def __call__(self, X1, X2=None):
    return self.evaluate(X1, X2)

  def evaluate(self, X1, X2=None):
    """ n1 x n2 Gram matrix; zeros((n1, n2)) if either side is empty (kernel.py:76-83). """
    X2 = X1 if X2 is None else X2
    if len(X1) == 0 or len(X2) == 0:
      return np.zeros((len(X1), len(X2)))
    return self._child_evaluate(X1, X2)

  def _child_evaluate(self, X1, X2):
    from .device import kernel_matrix      # late import: device.py needs torch + the library
    return kernel_matrix(self, X1, X2)

  def set_hyperparams(self, **kwargs):
    self.hyperparams = kwargs

  def add_hyperparams(self, **kwargs):
    for key, value in kwargs.items():
      self.hyperparams[key] = value

  def __str__(self):
    return '%s:: %s' % (type(self), str(self.hyperparams))


class SEKernel(Kernel):
  """ kernel.py:130-181: scale * exp(-||(x - y) / bw||^2 / 2). """

  def __init__(self, dim, scale=None, dim_bandwidths=None):
    super(SEKernel, self).__init__()
    self.dim = dim
    self.set_se_hyperparams(scale, dim_bandwidths)

  def is_guaranteed_psd(self):
    return True

  def set_dim_bandwidths(self, dim_bandwidths):
    if dim_bandwidths is not None:
      if len(dim_bandwidths) != self.dim:
        raise ValueError('Dimension of dim_bandwidths should be the same as dimension.')
      dim_bandwidths = np.array(dim_bandwidths).T
    self.add_hyperparams(dim_bandwidths=dim_bandwidths)

  def set_single_bandwidth(self, bandwidth):
    self.set_dim_bandwidths(None if bandwidth is None else [bandwidth] * self.dim)

  def set_scale(self, scale):
    self.add_hyperparams(scale=scale)

  def set_se_hyperparams(self, scale, dim_bandwidths):
    self.set_scale(scale)
    if hasattr(dim_bandwidths, '__len__'):
      self.set_dim_bandwidths(dim_bandwidths)
    else:
      self.set_single_bandwidth(dim_bandwidths)

  def change_smoothness(self, factor):
    self.hyperparams['dim_bandwidths'] *= factor

  # host-side metadata helpers of the reference's SEKernel (kernel.py:179-190): no kernel evaluation involved
  def get_scaled_repr(self, X):
    return X / self.hyperparams['dim_bandwidths']

  def get_effective_norm(self, X, order=None, is_single=True):
    scaled_X = self.get_scaled_repr(X)
    if is_single:
      return np.linalg.norm(scaled_X, ord=order)
    return np.array([np.linalg.norm(sx, ord=order) for sx in scaled_X])

  def __str__(self):
    return 'SE: sc:%0.4f avg-bw: %0.4f' % (self.hyperparams['scale'],
                                           np.mean(self.hyperparams['dim_bandwidths']))


def matern_constants(nu):
  """ The scalar constants of the half-integer Matern kernel formed exactly as the reference forms
      them (set_matern_hyperparams kernel.py:242-253, _eval_kernel_values_unnormalised :259-270). """
  if nu % 1 != 0.5:
    raise ValueError('Matern kernel: nu has to be p + 0.5 where p is an integer.')
  p = int(nu)
  if p > _lib.DFB_MAX_MATERN_P:
    raise NotImplementedError('Matern nu=%s: only nu <= %d.5 is supported on device.' % (
        nu, _lib.DFB_MAX_MATERN_P))
  coeffs = [math.factorial(p + i) / (math.factorial(i) * math.factorial(p - i))
            for i in range(p + 1)]
  gamma_ratio = math.gamma(p + 1) / math.gamma(2 * p + 1)
  s8 = float(np.sqrt(8 * nu))
  s2 = float(np.sqrt(2 * nu))
  u0 = 0
  for i in range(p + 1):
    u0 += coeffs[i] * (s8 * 0) ** (p - i)
  u0 *= (gamma_ratio * np.exp(-s2 * 0))
  return dict(p=p, s8=s8, s2=s2, coeffs=[float(c) for c in coeffs],
              gamma_ratio=float(gamma_ratio), norm_constant=float(1.0 / u0))


class MaternKernel(Kernel):
  """ kernel.py:224-299: half-integer Matern, nu = p + 1/2. """

  def __init__(self, dim, nu=None, scale=None, dim_bandwidths=None):
    super(MaternKernel, self).__init__()
    self.dim = dim
    self.p = None
    self.norm_constant = None
    self.set_matern_hyperparams(nu, scale, dim_bandwidths)

  def is_guaranteed_psd(self):
    return True

  def set_matern_hyperparams(self, nu, scale, dim_bandwidths):
    consts = matern_constants(nu)
    self.add_hyperparams(nu=nu)
    self.add_hyperparams(scale=scale)
    dim_bandwidths = dim_bandwidths if hasattr(dim_bandwidths, '__len__') else \
                     [dim_bandwidths] * self.dim
    self.add_hyperparams(dim_bandwidths=np.array(dim_bandwidths).T)
    self.p = consts['p']
    self.norm_constant = consts['norm_constant']

  def __str__(self):
    return 'Matern: nu=%0.1f sc:%0.4f avg-bw: %0.4f' % (
        self.hyperparams['nu'], self.hyperparams['scale'],
        np.mean(self.hyperparams['dim_bandwidths']))


class AdditiveKernel(Kernel):
  """ kernel.py:461-500: scale * sum_g k_g(x[g], y[g]) over non-overlapping groups. """

  def __init__(self, scale, kernel_list, groupings):
    if len(kernel_list) != len(groupings):
      raise ValueError('number of kernels do not correspond to number of groups.')
    super(AdditiveKernel, self).__init__()
    self.kernel_list = kernel_list
    self.groupings = groupings
    self.add_hyperparams(scale=scale)
    self.dim = sum([kern.dim for kern in self.kernel_list])

  def is_guaranteed_psd(self):
    return all([kern.is_guaranteed_psd() for kern in self.kernel_list])

  def __str__(self):
    return 'ADD scale=%0.2f, ' % (self.hyperparams['scale']) + ', '.join(
        ['%s(%s)' % (g, k) for (g, k) in zip(self.groupings, self.kernel_list)])


class CoordinateProductKernel(Kernel):
  """ kernel.py:541-590: scale * prod_i k_i(x[c_i], y[c_i]); the multi-fidelity kernel is the
      product of a fidelity-space and a domain kernel (euclidean_gp.py:369-374). """

  def __init__(self, dim, scale, kernel_list=None, coordinate_list=None):
    super(CoordinateProductKernel, self).__init__()
    self.dim = dim
    self.add_hyperparams(scale=scale)
    self.kernel_list = kernel_list
    self.coordinate_list = coordinate_list

  def set_kernel_list(self, kernel_list):
    self.kernel_list = kernel_list

  def set_new_kernel(self, kernel_idx, new_kernel):
    self.kernel_list[kernel_idx] = new_kernel

  def set_kernel_hyperparams(self, kernel_idx, **kwargs):
    self.kernel_list[kernel_idx].set_hyperparams(**kwargs)

  def is_guaranteed_psd(self):
    return all([kern.is_guaranteed_psd() for kern in self.kernel_list])

  def __str__(self):
    return 'CoordProd scale=%0.2f, ' % (self.hyperparams['scale']) + ', '.join(
        ['%s(%s)' % (g, k) for (g, k) in zip(self.coordinate_list, self.kernel_list)])


def kernel_from_spec(spec):
  """ Builds a kernel object from the nested-dict form used by synth_data.make_workload. """
  t = spec['type']
  if t == 'se':
    return SEKernel(spec['dim'], spec['scale'], spec['dim_bandwidths'])
  if t == 'matern':
    return MaternKernel(spec['dim'], spec['nu'], spec['scale'], spec['dim_bandwidths'])
  if t == 'additive':
    return AdditiveKernel(spec['scale'], [kernel_from_spec(s) for s in spec['kernels']],
                          spec['groupings'])
  if t == 'coordinate_product':
    return CoordinateProductKernel(spec['dim'], spec['scale'],
                                   [kernel_from_spec(s) for s in spec['kernels']],
                                   spec['coordinate_list'])
  raise ValueError('unknown kernel spec type %s' % (t))


# ---------------------------------------------------------------------------------------------
# Kernel object -> canonical sum-of-products form -> dfb_kernel_desc
# ---------------------------------------------------------------------------------------------
class _Factor(object):
  __slots__ = ('kind', 'p', 'scale', 's8', 's2', 'gamma_ratio', 'coeffs', 'train_coords',
               'cand_coords', 'bandwidths')


def _kind_of(kern):
  """ Duck-typed dispatch on the class name so the reference's own kernel objects work too. """
  names = [c.__name__ for c in type(kern).__mro__]
  for n in ('SEKernel', 'MaternKernel', 'AdditiveKernel', 'CoordinateProductKernel'):
    if n in names:
      return n
  raise NotImplementedError(
      'Kernel type %s is outside the B200 hot-path scope (supported: SEKernel, MaternKernel, '
      'AdditiveKernel, CoordinateProductKernel); there is no CPU fallback.' % (type(kern).__name__))


def _expand(kern, train_coords, cand_coords):
  """ Returns (post_scale, [(pre_scale, [factor, ...]), ...]) for `kern` applied to the given
      columns of the training / candidate matrices. """
  kind = _kind_of(kern)
  if kind in ('SEKernel', 'MaternKernel'):
    bws = np.asarray(kern.hyperparams['dim_bandwidths'], dtype=np.float64).reshape(-1)
    if len(bws) != len(train_coords):
      raise ValueError('kernel has %d bandwidths for %d coordinates' % (len(bws), len(train_coords)))
    f = _Factor()
    f.train_coords = [int(c) for c in train_coords]
    f.cand_coords = [int(c) for c in cand_coords]
    f.bandwidths = [float(b) for b in bws]
    if kind == 'SEKernel':
      f.kind = _lib.DFB_BASE_SE
      f.p, f.s8, f.s2, f.gamma_ratio, f.coeffs = 0, 0.0, 0.0, 0.0, []
      f.scale = float(kern.hyperparams['scale'])
    else:
      consts = matern_constants(kern.hyperparams['nu'])
      f.kind = _lib.DFB_BASE_MATERN
      f.p, f.s8, f.s2 = consts['p'], consts['s8'], consts['s2']
      f.gamma_ratio, f.coeffs = consts['gamma_ratio'], consts['coeffs']
      # K = hyperparams['scale'] * norm_constant * unnorm: the first product is formed on the host
      f.scale = float(kern.hyperparams['scale'] * consts['norm_constant'])
    return 1.0, [(1.0, [f])]
  if kind == 'AdditiveKernel':
    terms = []
    for sub, grp in zip(kern.kernel_list, kern.groupings):
      post, sub_terms = _expand(sub, [train_coords[g] for g in grp], [cand_coords[g] for g in grp])
      for pre, facs in sub_terms:
        terms.append((pre * post if post != 1.0 else pre, facs))
    return float(kern.hyperparams['scale']), terms
  # CoordinateProductKernel: distribute the product over the (usually single-term) children
  terms = [(float(kern.hyperparams['scale']), [])]
  for sub, crd in zip(kern.kernel_list, kern.coordinate_list):
    post, sub_terms = _expand(sub, [train_coords[c] for c in crd], [cand_coords[c] for c in crd])
    new_terms = []
    for pre_a, facs_a in terms:
      for pre_b, facs_b in sub_terms:
        scale_b = pre_b * post
        new_terms.append((pre_a * scale_b if scale_b != 1.0 else pre_a, facs_a + facs_b))
    terms = new_terms
  return 1.0, terms


def _base_at_zero(f):
  """ Base-kernel value at distance 0 in the device's operation order. """
  if f.kind == _lib.DFB_BASE_SE:
    return f.scale * np.exp(-0.0)
  u = 0.0
  for i in range(f.p + 1):
    e = f.p - i
    u = u + f.coeffs[i] * (1.0 if e == 0 else 0.0)
  u = u * (f.gamma_ratio * np.exp(-0.0))
  return f.scale * u


def build_descriptor(kern, train_dim=None, cand_coords=None, train_coords=None, cand_dim=None):
  """ kern -> _lib.KernelDesc.  By default train and candidate matrices share the column layout
      (kernel applied to columns 0..dim-1).  Add-UCB passes train_coords = the group's columns of
      the training matrix and cand_coords = 0..d_j-1 (gpb_acquisitions.py:160-168). """
  dim = int(kern.dim)
  if train_coords is None:
    train_coords = list(range(dim))
  if cand_coords is None:
    cand_coords = list(range(dim))
  if train_dim is None:
    train_dim = max(train_coords) + 1
  if cand_dim is None:
    cand_dim = max(cand_coords) + 1
  post, terms = _expand(kern, train_coords, cand_coords)
  n_factors = sum(len(facs) for _, facs in terms)
  n_slots = sum(len(f.bandwidths) for _, facs in terms for f in facs)
  if len(terms) > _lib.DFB_MAX_TERMS or n_factors > _lib.DFB_MAX_FACTORS or \
     n_slots > _lib.DFB_MAX_SLOTS:
    raise NotImplementedError('kernel too large for the device descriptor: %d terms, %d factors, '
                              '%d slots' % (len(terms), n_factors, n_slots))
  d = _lib.KernelDesc()
  d.n_terms, d.n_factors, d.n_slots = len(terms), n_factors, n_slots
  d.train_dim, d.cand_dim = int(train_dim), int(cand_dim)
  d.post_scale = float(post)
  fi, si = 0, 0
  total = 0.0
  for ti, (pre, facs) in enumerate(terms):
    d.term_first_factor[ti] = fi
    d.term_pre_scale[ti] = float(pre)
    prod = float(pre)
    for f in facs:
      fd = d.factors[fi]
      fd.kind, fd.p, fd.n_dims, fd.slot_off = f.kind, f.p, len(f.bandwidths), si
      fd.scale, fd.s8, fd.s2, fd.gamma_ratio = f.scale, f.s8, f.s2, f.gamma_ratio
      for i, cval in enumerate(f.coeffs):
        fd.coeffs[i] = cval
      for q in range(len(f.bandwidths)):
        d.slot_train_coord[si] = f.train_coords[q]
        d.slot_cand_coord[si] = f.cand_coords[q]
        d.slot_bandwidth[si] = f.bandwidths[q]
        si += 1
      prod = prod * _base_at_zero(f)
      fi += 1
    total = total + prod
  d.term_first_factor[len(terms)] = fi
  d.kss = float(post * total)
  return d
