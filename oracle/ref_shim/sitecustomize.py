"""NumPy-2 alias shim so the UNMODIFIED reference (dragonfly-opt 0.1.7) imports in this image.

Test infrastructure only (see oracle/__init__.py).  Restores aliases removed from NumPy>=1.24/2.0
that the reference still uses on the hot path: np.math (kernel.py:263-268), np.asscalar
(kernel.py:734), np.object (general_utils.py:139), np.int (oper_utils.py:340).
"""
import math
import numpy as np

for _k, _v in dict(math=math, object=object, int=int, float=float, bool=bool).items():
  if not hasattr(np, _k):
    setattr(np, _k, _v)
if not hasattr(np, 'asscalar'):
  np.asscalar = lambda a: np.asarray(a).item()
