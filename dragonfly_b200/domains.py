"""
Minimal mirror of dragonfly/exd/domains.py:71-115 (EuclideanDomain) -- the only domain type on the
hot path.  anc_data.domain only needs get_type(), get_dim() and .bounds.
"""
import numpy as np


class EuclideanDomain(object):
  """ Domain for Euclidean spaces: bounds is a (dim, 2) array of [lower, upper]. """

  def __init__(self, bounds):
    self.bounds = np.array(bounds, dtype=np.float64)
    self.diameter = np.linalg.norm(self.bounds[:, 1] - self.bounds[:, 0])
    self.dim = len(bounds)

  def get_type(self):
    return 'euclidean'

  def get_dim(self):
    return self.dim

  def is_a_member(self, point):
    point = np.asarray(point)
    return bool(len(point) == self.dim and np.all(point >= self.bounds[:, 0])
                and np.all(point <= self.bounds[:, 1]))

  def __str__(self):
    return 'Euclidean: %s' % (self.bounds.tolist())
