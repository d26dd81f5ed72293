"""CPU baseline chunk-size sweep: the faithful chunked gp.eval(chunk, 'std') + EI + arg-max driver of
oracle/gp_oracle.py at N = 5000 for chunk in {500, 2000, 8000}, to show that bench.py's chunk of 2000 is not an
adversarial choice for the reference.  Runs on the host only.  Usage: python tests/cpu_chunk_sweep.py [n_cand]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dragonfly_b200 import synth_data  # noqa: E402
from oracle import gp_oracle as O      # noqa: E402  (lives under tests/: only tests/, smoke() and bench.py's CPU leg may import the oracle)

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=5000, n_cand=n_cand)
k = w['kernel']
gp = O.OGP(w['X'], w['Y'], O.OMaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
           lambda x: np.array([w['mean_const']] * len(x)), w['noise_var'])
out = {'n_train': 5000, 'candidates': n_cand, 'cores': os.cpu_count()}
for chunk in (500, 2000, 8000):
  t0 = time.perf_counter()
  val, idx, _ = O.chunked_scores(gp, w['candidates'], 'ei', chunk=chunk, curr_best=float(w['Y'].max()))
  dt = time.perf_counter() - t0
  out['chunk_%d' % chunk] = {'seconds': dt, 'cands_per_s': n_cand / dt, 'argmax': int(idx)}
  print(chunk, out['chunk_%d' % chunk], flush=True)
print(json.dumps(out))
