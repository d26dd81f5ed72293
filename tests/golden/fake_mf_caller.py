"""A deterministic stand-in for the multi-fidelity function caller BOCA consults
(experiment_caller.py:496-528): candidate fidelities, cost ratios and information gaps.
Used by make_golden.py (against the reference) and by the tests (against the device path)."""
import numpy as np


class FakeMFCaller(object):

  def __init__(self, fidel_to_opt):
    self.fidel_to_opt = np.asarray(fidel_to_opt, dtype=np.float64)

  def get_candidate_fidels_and_cost_ratios(self, point, filter_by_cost=True):
    fidels = [np.array([z]) for z in np.linspace(0.05, 0.95, 19)]
    ratios = [float((0.1 + 0.9 * f[0] ** 2) / 1.0) for f in fidels]
    return fidels, ratios

  def get_information_gap(self, fidels):
    return [float(np.linalg.norm(np.asarray(f) - self.fidel_to_opt)) for f in fidels]
