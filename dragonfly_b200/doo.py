"""
Parallel deterministic optimistic optimisation (PDOO) of an acquisition, with the children of every split scored
in ONE batched call -- SURVEY.md 8f rank 4.

What it replaces: dragonfly/utils/doo.py (OptTree.run_PDOO / run_DOO / split_children / querie, pdoo_wrap) as
driven by pdoo_maximise (dragonfly/utils/oper_utils.py:257-271: K = 2, tol = 1e-3, nu_max = 1, C_init = 0.8,
rho_max = 0.9, POO_mult = 0.5) -- the maximiser the reference ends up in for acq_opt_method 'pdoo', and for the
default 'direct' whenever its Fortran DIRECT extension is not built (oper_utils.py:121-137).  The reference
evaluates the acquisition one point per Python call (gpb_acquisitions.py:33-37, doo.py:117-125); here the K
children of a split go to the device together (GP.eval's row-streaming path for <= 16 points), which is the
only change: the search itself -- cell geometry, optimistic bounds, fidelity bookkeeping, evaluation cache,
priority order, budget accounting, the final selection -- is restated so that, for the same objective values, it
visits the same cells in the same order and returns the same point (tests/test_host_logic.py pins it on the
reference's own runs, tests/golden/pdoo.npz).

The objective here never depends on the "fidelity" z (cost 1 per evaluation), but z still decides when a cached
cell is re-evaluated (|z_old - z| > tol) and enters the bounds, so it is carried exactly as in the reference.
"""
import heapq

import numpy as np


class _Leaf(object):
  """ A cell with its value and optimistic bound; ordered for heapq so that the LARGEST bound pops first and
      equal bounds compare equal (doo.py:84-91). """
  __slots__ = ('cell', 'value', 'fidelity', 'bound', 'height', 'split_dim')

  def __init__(self, cell, value, fidelity, bound, height, split_dim):
    self.cell, self.value, self.fidelity = cell, value, fidelity
    self.bound, self.height, self.split_dim = bound, height, split_dim

  def __lt__(self, other):
    return other.bound < self.bound

  def __eq__(self, other):
    return other.bound == self.bound


class PDOOSearch(object):
  """ One PDOO run over the unit cube; `batch_obj(P)` maps an (k, d) array of NORMALISED points to k values. """

  def __init__(self, batch_obj, dim, total_budget, nu_max=1.0, rho_max=0.9, K=2, C_init=0.8, tol=1e-3):
    self.batch_obj, self.dim = batch_obj, dim
    self.total_budget, self.nu_max, self.rho_max, self.K = total_budget, nu_max, rho_max, K
    self.C, self.tol = C_init, tol
    self.cache = {}                  # cell -> _Leaf holding the last value / fidelity seen for it
    self.query_pts, self.query_vals = [], []
    self.num_device_calls = 0

  # -- evaluation of a group of cells (one split's children, or the root) ---------------------------------------
  def _fidelity(self, diam):
    return min(max(1 - diam / self.C, self.tol), 1.0)

  def _score_cells(self, cells, height, rho, nu, split_dim):
    """ doo.py:127-158 for each cell IN ORDER, with the needed objective values fetched in one batched call.
        Which cells need a value does not depend on the values themselves unless C doubles mid-group (it cannot
        for an objective that ignores z: a re-evaluation returns the cached value); that case falls back to
        one-by-one evaluation to stay exact. """
    diam = nu * (rho ** height)
    C_before = self.C
    z = self._fidelity(diam)
    need = [c for c in cells if not (c in self.cache and abs(self.cache[c].fidelity - z) <= self.tol)]
    values = {}
    if need:
      pts = np.array([[(lo + hi) / 2.0 for (lo, hi) in c] for c in need])
      got = np.asarray(self.batch_obj(pts), dtype=np.float64).reshape(-1)
      self.num_device_calls += 1
      values = dict(zip(need, [float(v) for v in got]))
    leaves, cost = [], 0
    for c in cells:
      if self.C != C_before:         # C doubled inside this group: the remaining cells see a new z
        z = self._fidelity(diam)
        if c not in values and not (c in self.cache and abs(self.cache[c].fidelity - z) <= self.tol):
          mid = np.array([[(lo + hi) / 2.0 for (lo, hi) in c]])
          values[c] = float(np.asarray(self.batch_obj(mid)).reshape(-1)[0])
          self.num_device_calls += 1
      if c in self.cache:
        known = self.cache[c]
        if abs(known.fidelity - z) <= self.tol:
          value, spent = known.value, 0
        else:
          value = self._record(c, values[c])
          if abs(value - known.value) > self.C * abs(known.fidelity - z):
            self.C = 2.0 * self.C
          known.value, known.fidelity = value, z
          spent = 1.0
      else:
        value = self._record(c, values[c])
        self.cache[c] = _Leaf(c, value, z, diam + self.C * (1.0 - z) + value, height, split_dim)
        spent = 1.0
      leaves.append(_Leaf(c, value, z, diam + self.C * (1.0 - z) + value, height, split_dim))
      cost += spent
    return leaves, cost

  def _record(self, cell, value):
    if len(self.query_vals) <= self.total_budget:
      self.query_pts.append(np.array([(lo + hi) / 2.0 for (lo, hi) in cell]))
      self.query_vals.append(value)
    return value

  # -- one DOO pass (doo.py:188-230) -----------------------------------------------------------------------------------
  def _split(self, leaf, rho, nu):
    spans = [abs(hi - lo) for (lo, hi) in leaf.cell]
    d = int(np.argmax(spans))
    if d == leaf.split_dim:
      d = (leaf.split_dim - 1) % len(leaf.cell)
    edges = np.linspace(leaf.cell[d][0], leaf.cell[d][1], self.K + 1)
    kids = [tuple((edges[i], edges[i + 1]) if j == d else side for j, side in enumerate(leaf.cell))
            for i in range(self.K)]
    return self._score_cells(kids, leaf.height + 1, rho, nu, d)

  def _doo_pass(self, budget, nu, rho):
    root = tuple((0, 1) for _ in range(self.dim))
    leaves, cost = self._score_cells([root], 0, rho, nu, 0)
    heap = []
    heapq.heappush(heap, leaves[0])
    seen = {}
    while cost <= budget:
      top = heapq.heappop(heap)
      seen[top.cell] = (top.value, top.fidelity, top.height)
      kids, spent = self._split(top, rho, nu)
      if top.cell == kids[0].cell:
        break
      cost = cost + spent
      for kid in kids:
        heapq.heappush(heap, kid)
    while heap:
      leaf = heapq.heappop(heap)
      seen[leaf.cell] = (leaf.value, leaf.fidelity, leaf.height)
    best, pick = float('-inf'), None
    for cell, (value, fidelity, height) in seen.items():
      if value - self.C * (1.0 - fidelity) > best:
        best = value - self.C * (1.0 - fidelity)
        pick = (value, fidelity, np.array([(lo + hi) / 2 for (lo, hi) in cell]), height)
    return pick[0], pick[1], pick[2], cost, pick[3]

  # -- the sweep over rho (doo.py:232-251) ------------------------------------------------------------------------------
  def run(self, mult=0.5):
    Dm = int(np.log(self.K) / np.log(1 / self.rho_max))
    n = self.total_budget / 1.0
    N = int(mult * Dm * np.log(n / np.log(n)))
    budget = self.total_budget / float(N)
    passes = []
    for i in range(N):
      rho = (self.rho_max) ** (float(N) / (N - i))
      passes.append(self._doo_pass(budget, self.nu_max, rho))
    adjusted = [p[0] - self.C * (1 - p[1]) for p in passes]
    return passes, int(np.argmax(adjusted))


def pdoo_maximise(obj, bounds, max_evals, vectorised=True):
  """ oper_utils.py:257-271 + doo.py:253-260: returns (max_val, max_pt, None).  `obj` takes an (k, d) array of
      points in the ORIGINAL coordinates and returns k values when `vectorised` (the acquisition closures of
      gpb_acquisitions do); otherwise it is called one point at a time like the reference does. """
  bounds = np.array(bounds)
  lo, width = bounds[:, 0], bounds[:, 1] - bounds[:, 0]

  def batch_obj(P):
    X = P * width + lo                                         # map_to_bounds (general_utils.py:25-27)
    if vectorised:
      return obj(X)
    return np.array([float(obj(x)) for x in X])
  search = PDOOSearch(batch_obj, len(bounds), float(max_evals), nu_max=1.0, rho_max=0.9, K=2, C_init=0.8,
                      tol=1e-3)
  passes, index = search.run(0.5)
  max_pt = passes[index][2] * width + lo
  pdoo_maximise.last_search = search            # diagnostics: query sequence, number of batched calls
  return passes[index][0], max_pt, None


pdoo_maximise.last_search = None
