"""
Device-side objects behind the Python host: a `DevicePosterior` wraps one libdfb200 handle plus the
torch-owned workspace it computes in.  PyTorch appears here only as the allocator of device buffers
and the owner of CUDA streams; every numeric result comes from libdfb200's CUDA kernels.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from .kernel import build_descriptor


def _require_cuda(device=None):
  if not torch.cuda.is_available():
    raise RuntimeError('dragonfly_b200 needs a CUDA device (B200, sm_100a); none is visible and '
                       'there is no CPU fallback.')
  if device is None:
    device = torch.cuda.current_device()
  return torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)


def _dev_f64(arr, device):
  """ Host array-like or torch tensor -> contiguous fp64 CUDA tensor on `device`. """
  if isinstance(arr, torch.Tensor):
    return arr.to(device=device, dtype=torch.float64).contiguous()
  return torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=np.float64))).to(device)


# Options applied to every new DevicePosterior (name -> int), e.g. {'score_impl': 0} to force the fp64
# DMMA contraction everywhere.  See dfb_set_option in include/dfb200.h.
DEFAULT_OPTIONS = {}

# A BO loop builds a fresh GP (hence a fresh handle) every iteration; the ~1.3 GB workspace of an N = 5000
# posterior is recycled through this small per-(device, size) pool rather than through cudaMalloc /
# cudaFree, whose cost (tens of ms, variable) would otherwise sit inside every build_posterior.
_WORKSPACE_POOL = {}
_WORKSPACE_POOL_DEPTH = 2


def _take_workspace(key, dev):
  free = _WORKSPACE_POOL.get(key)
  if free:
    return free.pop()
  return torch.empty(key[1] + 256, dtype=torch.uint8, device=dev)


def _give_workspace(key, ws):
  if ws is None:
    return
  free = _WORKSPACE_POOL.setdefault(key, [])
  if len(free) < _WORKSPACE_POOL_DEPTH:
    free.append(ws)


def release_workspaces():
  """ Drops the pooled workspaces (returns the memory to torch's allocator). """
  _WORKSPACE_POOL.clear()


class DevicePosterior(object):
  """ One handle + workspace.  Immutable once built (GP objects replace, never mutate, it), so
      shallow / deep copies of a GP may share it (gpb_acquisitions.py:104, unittest_mf_gp.py:109). """

  def __init__(self, n_max, device=None, chunk=0):
    self.lib = _lib.load()
    self.device = _require_cuda(device)
    self.n_max = int(n_max)
    hp = C.c_void_p()
    _lib.check(self.lib.dfb_create(C.byref(hp), self.device.index), 'dfb_create')
    self.h = hp
    nbytes = self.lib.dfb_workspace_bytes(self.n_max, 0, int(chunk))
    self._pool_key = (self.device.index, int(nbytes))
    self.workspace = _take_workspace(self._pool_key, self.device)
    ptr = (self.workspace.data_ptr() + 255) // 256 * 256
    with torch.cuda.device(self.device):
      stream = torch.cuda.current_stream(self.device).cuda_stream
      _lib.check(self.lib.dfb_set_stream(self.h, C.c_void_p(stream)), 'dfb_set_stream')
      _lib.check(self.lib.dfb_set_workspace(self.h, C.c_void_p(ptr), C.c_size_t(nbytes), self.n_max,
                                            int(chunk)), 'dfb_set_workspace')
    self.n = 0
    self.dim = 0
    self.lml = None
    self._owners = weakref.WeakSet()      # GP objects using this posterior (gp_core.GP._post_is_shared)
    # joint-posterior blocks (covariance / Thompson draws) must fit the handle's scoring chunk
    self.TS_BLOCK = min(DevicePosterior.TS_BLOCK, int(self.query('chunk')))
    for _name, _value in DEFAULT_OPTIONS.items():
      self.set_option(_name, _value)
    self._keep = []      # tensors that must outlive asynchronous use

  def __del__(self):
    try:
      if getattr(self, 'h', None) is not None and self.h.value:
        torch.cuda.synchronize(self.device)
        self.lib.dfb_destroy(self.h)
        self.h = None
        _give_workspace(self._pool_key, self.workspace)
        self.workspace = None
    except Exception:  # pylint: disable=broad-except
      pass

  def bind_current_stream(self):
    """ Issue this handle's work on the calling thread's current torch stream from now on. """
    stream = torch.cuda.current_stream(self.device).cuda_stream
    _lib.check(self.lib.dfb_set_stream(self.h, C.c_void_p(stream)), 'dfb_set_stream')

  # -- model ------------------------------------------------------------------------------------
  def set_kernel(self, desc):
    _lib.check(self.lib.dfb_set_kernel(self.h, C.byref(desc)), 'dfb_set_kernel')

  def set_test_kernel(self, desc):
    if desc is None:
      _lib.check(self.lib.dfb_set_test_kernel(self.h, None), 'dfb_set_test_kernel')
    else:
      _lib.check(self.lib.dfb_set_test_kernel(self.h, C.byref(desc)), 'dfb_set_test_kernel')

  def set_train(self, X, y_centred):
    Xd = _dev_f64(X, self.device)
    yd = _dev_f64(y_centred, self.device)
    self.n, self.dim = int(Xd.shape[0]), int(Xd.shape[1])
    _lib.check(self.lib.dfb_set_train(self.h, C.c_void_p(Xd.data_ptr()), self.n, self.dim,
                                      C.c_void_p(yd.data_ptr())), 'dfb_set_train')
    torch.cuda.synchronize(self.device)

  def build(self, noise_var, jitter=0.0, flags=_lib.DFB_BUILD_FULL):
    """ Returns (info, lml): info > 0 means 'not positive definite' (np.linalg.LinAlgError). """
    lml = C.c_double(0.0)
    info = _lib.check(self.lib.dfb_build_posterior(self.h, float(noise_var), float(jitter), int(flags),
                                                   C.byref(lml)), 'dfb_build_posterior')
    self.lml = lml.value if info == 0 else None
    return info, self.lml

  def capacity(self):
    """ Training points this posterior can hold without a rebuild: its padded size (a multiple of 128). """
    return (self.n + 127) // 128 * 128

  def extend(self, X_new, y_centred_new, flags=_lib.DFB_BUILD_FULL, save=False):
    """ dfb_extend_posterior: appends training points to the built posterior in O(N^2) work.
        Returns (info, lml); info > 0 = the extended matrix is not positive definite (with save=True
        the un-extended posterior is back in place, otherwise it must be rebuilt). """
    Xd = _dev_f64(X_new, self.device)
    yd = _dev_f64(y_centred_new, self.device)
    q = int(Xd.shape[0])
    assert int(Xd.shape[1]) == self.dim and int(yd.shape[0]) == q
    lml = C.c_double(0.0)
    info = _lib.check(self.lib.dfb_extend_posterior(
        self.h, C.c_void_p(Xd.data_ptr()), q, C.c_void_p(yd.data_ptr()),
        int(flags) | (_lib.DFB_EXTEND_SAVE if save else 0), C.byref(lml)), 'dfb_extend_posterior')
    if info == 0:
      self.n += q
      self.lml = lml.value
    return info, (lml.value if info == 0 else None)

  def restore(self, n_before):
    """ dfb_restore_posterior: undoes an extend(..., save=True) bit for bit. """
    _lib.check(self.lib.dfb_restore_posterior(self.h), 'dfb_restore_posterior')
    self.n = int(n_before)

  def lml_gradients(self, dim):
    """ dfb_lml_gradients: [scale, noise_var / noise_var, noise_mean, same_dim_bandwidths, dim_bandwidths[0..d)]. """
    out = (C.c_double * (4 + int(dim)))()
    st = self.lib.dfb_lml_gradients(self.h, out, 4 + int(dim))
    if st == -3:
      raise NotImplementedError(_lib.last_error())
    _lib.check(st, 'dfb_lml_gradients')
    return np.array(out[:], dtype=np.float64)

  def max_diag(self):
    out = C.c_double(0.0)
    _lib.check(self.lib.dfb_get_max_diag(self.h, C.byref(out)), 'dfb_get_max_diag')
    return out.value

  def get_state(self, want_L=False, want_alpha=False, want_K=False):
    L = torch.empty((self.n, self.n), dtype=torch.float64, device=self.device) if want_L else None
    a = torch.empty((self.n,), dtype=torch.float64, device=self.device) if want_alpha else None
    K = torch.empty((self.n, self.n), dtype=torch.float64, device=self.device) if want_K else None
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(self.lib.dfb_get_state(self.h, ptr(L), ptr(a), ptr(K)), 'dfb_get_state')
    return L, a, K

  def set_alpha(self, alpha):
    ad = _dev_f64(alpha, self.device)
    _lib.check(self.lib.dfb_set_alpha(self.h, C.c_void_p(ad.data_ptr()), int(ad.shape[0])),
               'dfb_set_alpha')
    torch.cuda.synchronize(self.device)

  # -- prediction ---------------------------------------------------------------------------------
  def eval(self, Xc, mean_const=0.0, want_std=True):
    """ Xc: host ndarray (results are host ndarrays, copies inside the C call) or CUDA tensor
        (results are CUDA tensors).  Returns (mu, sd or None). """
    if isinstance(Xc, torch.Tensor):
      Xd = _dev_f64(Xc, self.device)
      m, dc = int(Xd.shape[0]), int(Xd.shape[1])
      mu = torch.empty((m,), dtype=torch.float64, device=self.device)
      sd = torch.empty((m,), dtype=torch.float64, device=self.device) if want_std else None
      _lib.check(self.lib.dfb_eval(self.h, C.c_void_p(Xd.data_ptr()), m, dc, _lib.DFB_DEVICE,
                                   float(mean_const), C.c_void_p(mu.data_ptr()),
                                   C.c_void_p(sd.data_ptr()) if want_std else None), 'dfb_eval')
      return mu, sd
    Xh = np.ascontiguousarray(np.asarray(Xc, dtype=np.float64))
    m, dc = Xh.shape
    mu = np.empty((m,), dtype=np.float64)
    sd = np.empty((m,), dtype=np.float64) if want_std else None
    _lib.check(self.lib.dfb_eval(self.h, Xh.ctypes.data_as(C.c_void_p), m, dc, _lib.DFB_HOST,
                                 float(mean_const), mu.ctypes.data_as(C.c_void_p),
                                 sd.ctypes.data_as(C.c_void_p) if want_std else None), 'dfb_eval')
    return mu, sd

  def score_argmax(self, acq_desc, Xc, mean_const=0.0, want_scores=False):
    """ Fused scoring + arg-max.  Returns (best_score, best_index, scores or None). """
    bs, bi = C.c_double(0.0), C.c_int64(-1)
    if isinstance(Xc, torch.Tensor):
      Xd = _dev_f64(Xc, self.device)
      m, dc = int(Xd.shape[0]), int(Xd.shape[1])
      sc = torch.empty((m,), dtype=torch.float64, device=self.device) if want_scores else None
      _lib.check(self.lib.dfb_score_argmax(
          self.h, C.byref(acq_desc), C.c_void_p(Xd.data_ptr()), m, dc, _lib.DFB_DEVICE,
          float(mean_const), C.c_void_p(sc.data_ptr()) if want_scores else None, C.byref(bs),
          C.byref(bi)), 'dfb_score_argmax')
      return bs.value, bi.value, sc
    Xh = np.ascontiguousarray(np.asarray(Xc, dtype=np.float64))
    m, dc = Xh.shape
    sc = np.empty((m,), dtype=np.float64) if want_scores else None
    _lib.check(self.lib.dfb_score_argmax(
        self.h, C.byref(acq_desc), Xh.ctypes.data_as(C.c_void_p), m, dc, _lib.DFB_HOST,
        float(mean_const), sc.ctypes.data_as(C.c_void_p) if want_scores else None, C.byref(bs),
        C.byref(bi)), 'dfb_score_argmax')
    return bs.value, bi.value, sc

  # -- joint posterior over one block: covariance and Thompson draws -----------------------------
  TS_BLOCK = 4096

  def _ensure_ts_workspace(self, m):
    mb = getattr(self, '_ts_mb', 0)
    if mb >= m:
      return
    mb = max(int(m), 512)
    nbytes = self.lib.dfb_ts_workspace_bytes(self.n_max, mb)
    self._ts_workspace = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
    ptr = (self._ts_workspace.data_ptr() + 255) // 256 * 256
    _lib.check(self.lib.dfb_set_ts_workspace(self.h, C.c_void_p(ptr), C.c_size_t(nbytes), mb),
               'dfb_set_ts_workspace')
    self._ts_mb = mb

  def eval_covar(self, Xc, mean_const=0.0):
    """ (mu, covar) of GP.eval(X, 'covar') for one block; host ndarray in -> host ndarrays out. """
    Xd = _dev_f64(Xc, self.device)
    m, dc = int(Xd.shape[0]), int(Xd.shape[1])
    self._ensure_ts_workspace(m)
    mu = torch.empty((m,), dtype=torch.float64, device=self.device)
    cov = torch.empty((m, m), dtype=torch.float64, device=self.device)
    _lib.check(self.lib.dfb_eval_covar(self.h, C.c_void_p(Xd.data_ptr()), m, dc, float(mean_const),
                                       C.c_void_p(mu.data_ptr()), C.c_void_p(cov.data_ptr())),
               'dfb_eval_covar')
    if isinstance(Xc, torch.Tensor):
      return mu, cov
    return mu.cpu().numpy(), cov.cpu().numpy()

  def ts_draws(self, Xc, Ut, mean_const=0.0, jitter=0.0):
    """ One attempt of draw_gaussian_samples on a block: returns (info, samples (S, m), max_diag). """
    Xd = _dev_f64(Xc, self.device)
    Ud = _dev_f64(Ut, self.device)
    m, dc = int(Xd.shape[0]), int(Xd.shape[1])
    S = int(Ud.shape[0])
    assert int(Ud.shape[1]) == m
    self._ensure_ts_workspace(m)
    out = torch.empty((S, m), dtype=torch.float64, device=self.device)
    mx = C.c_double(0.0)
    info = _lib.check(self.lib.dfb_ts_draws(self.h, C.c_void_p(Xd.data_ptr()), m, dc, float(mean_const),
                                            C.c_void_p(Ud.data_ptr()), S, float(jitter),
                                            C.c_void_p(out.data_ptr()), None, C.byref(mx)),
                      'dfb_ts_draws')
    return info, out, mx.value

  def moo_score_argmax(self, kind, a_list, b_list, weights, refs=None, beta=0.0, want_scores=False):
    """ dfb_moo_score_argmax: scalarise n_obj objectives' device vectors and take the arg-max.
        a_list / b_list: CUDA fp64 tensors (mu_k or sampled values; sd_k or None).
        Returns (best_score, best_index, scores tensor or None). """
    K = len(a_list)
    d = _lib.MooDesc()
    d.kind, d.n_obj, d.beta = int(kind), K, float(beta)
    for k in range(K):
      d.weight[k] = float(weights[k])
      d.ref[k] = float(refs[k]) if refs is not None else 0.0
    a_t = [_dev_f64(a, self.device) for a in a_list]
    b_t = [_dev_f64(b, self.device) for b in b_list] if b_list is not None else None
    m = int(a_t[0].shape[0])
    assert all(int(t.shape[0]) == m for t in a_t) and (b_t is None or all(int(t.shape[0]) == m for t in b_t))
    PtrArr = C.c_void_p * K
    a_ptrs = PtrArr(*[t.data_ptr() for t in a_t])
    b_ptrs = PtrArr(*[t.data_ptr() for t in b_t]) if b_t is not None else None
    sc = torch.empty((m,), dtype=torch.float64, device=self.device) if want_scores else None
    bs, bi = C.c_double(0.0), C.c_int64(-1)
    _lib.check(self.lib.dfb_moo_score_argmax(
        self.h, C.byref(d), a_ptrs, b_ptrs, m, C.c_void_p(sc.data_ptr()) if want_scores else None,
        C.byref(bs), C.byref(bi)), 'dfb_moo_score_argmax')
    return bs.value, bi.value, sc

  def fill_rng(self, seed, col0, S, m, what=_lib.DFB_RNG_NORMAL, out=None):
    """ dfb_fill_rng: the S x m matrix of counter-based normals / uniforms for global columns col0 .. col0+m-1. """
    if out is None:
      out = torch.empty((int(S), int(m)), dtype=torch.float64, device=self.device)
    _lib.check(self.lib.dfb_fill_rng(self.h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), int(col0), int(S), int(m),
                                     int(what), C.c_void_p(out.data_ptr())), 'dfb_fill_rng')
    return out

  def fill_candidates(self, seed, row0, m, bounds, out=None):
    """ dfb_fill_candidates: rows row0 .. row0+m-1 of the device-generated candidate matrix (m x d CUDA tensor);
        bounds: (d, 2) array-like of [lo, hi]. """
    b = np.ascontiguousarray(np.asarray(bounds, dtype=np.float64))
    d = int(b.shape[0])
    lo = np.ascontiguousarray(b[:, 0]); hi = np.ascontiguousarray(b[:, 1])
    if out is None:
      out = torch.empty((int(m), d), dtype=torch.float64, device=self.device)
    _lib.check(self.lib.dfb_fill_candidates(self.h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), int(row0), int(m), d,
                                            lo.ctypes.data_as(C.POINTER(C.c_double)),
                                            hi.ctypes.data_as(C.POINTER(C.c_double)),
                                            C.c_void_p(out.data_ptr())), 'dfb_fill_candidates')
    return out

  def ts_argmax(self, samples, idx_base, best, index, reset):
    """ dfb_ts_argmax: fold one block of draws (S x m CUDA tensor) into the running per-draw arg-max. """
    S, m = int(samples.shape[0]), int(samples.shape[1])
    _lib.check(self.lib.dfb_ts_argmax(self.h, C.c_void_p(samples.data_ptr()), int(samples.stride(0)), S, m,
                                      int(idx_base), 1 if reset else 0, C.c_void_p(best.data_ptr()),
                                      C.c_void_p(index.data_ptr())), 'dfb_ts_argmax')

  def launch_count(self):
    return int(self.lib.dfb_launch_count(self.h))

  def set_option(self, name, value):
    _lib.check(self.lib.dfb_set_option(self.h, name.encode('utf-8'), int(value)), 'dfb_set_option')

  def query(self, name):
    out = C.c_double(0.0)
    _lib.check(self.lib.dfb_query(self.h, name.encode('utf-8'), C.byref(out)), 'dfb_query')
    return out.value

  def profile_enable(self, on=True):
    _lib.check(self.lib.dfb_profile_enable(self.h, 1 if on else 0), 'dfb_profile_enable')

  def profile_read(self, cls):
    """ (milliseconds, launches, work units) accumulated for a kernel class; resets it. """
    ms, n, u = C.c_double(0.0), C.c_int64(0), C.c_double(0.0)
    _lib.check(self.lib.dfb_profile_read(self.h, int(cls), C.byref(ms), C.byref(n), C.byref(u)),
               'dfb_profile_read')
    return ms.value, n.value, u.value


def measure_peak(what, device=None):
  """ dfb_measure_peak: 'i8' -> tcgen05 kind::i8 issue rate (int8 TOP/s), 'f64' -> DMMA issue rate (TFLOP/s). """
  dev = _require_cuda(device)
  out = C.c_double(0.0)
  code = {'i8': _lib.DFB_PEAK_TCGEN05_I8, 'f64': _lib.DFB_PEAK_DMMA_F64}[what]
  _lib.check(_lib.load().dfb_measure_peak(dev.index, code, C.byref(out)), 'dfb_measure_peak')
  return out.value


_KM_CACHE = {}


def kernel_matrix(kern, X1, X2, device=None):
  """ Kernel.__call__(X1, X2) on the GPU (dfb_kernel_matrix).  Returns a host ndarray (n1, n2). """
  dev = _require_cuda(device)
  X1d = _dev_f64(np.asarray(X1, dtype=np.float64), dev)
  X2d = _dev_f64(np.asarray(X2, dtype=np.float64), dev)
  n1, d1 = int(X1d.shape[0]), int(X1d.shape[1])
  n2, d2 = int(X2d.shape[0]), int(X2d.shape[1])
  key = (dev.index,)
  post = _KM_CACHE.get(key)
  if post is None or post.n_max < n2:
    post = DevicePosterior(max(n2, 1024), device=dev, chunk=128)
    _KM_CACHE[key] = post
  desc = build_descriptor(kern, train_dim=d1, cand_dim=d1)
  K = torch.empty((n1, n2), dtype=torch.float64, device=dev)
  _lib.check(post.lib.dfb_kernel_matrix(post.h, C.byref(desc), C.c_void_p(X1d.data_ptr()), n1, d1,
                                        C.c_void_p(X2d.data_ptr()), n2, d2,
                                        C.c_void_p(K.data_ptr())), 'dfb_kernel_matrix')
  return K.cpu().numpy()


def make_acq_desc(kind, beta=0.0, best=0.0, ref_mean=0.0, ref_std=0.0):
  a = _lib.AcqDesc()
  a.kind = {'mean': _lib.DFB_ACQ_MEAN, 'ucb': _lib.DFB_ACQ_UCB, 'ei': _lib.DFB_ACQ_EI,
            'pi': _lib.DFB_ACQ_PI, 'ttei': _lib.DFB_ACQ_TTEI}[kind]
  a.beta, a.best, a.ref_mean, a.ref_std = float(beta), float(best), float(ref_mean), float(ref_std)
  return a
