"""Scoring time of 10^6 candidates at the headline geometry for several scoring-chunk sizes (gp_core.DEFAULT_CHUNK)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, gp_core, device
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=5000, n_cand=16)
acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
cd = torch.from_numpy(np.random.RandomState(1000).random_sample((1000000, 6))).cuda()
for chunk in [0, 9856, 13056, 19712, 26112]:
  gp_core.DEFAULT_CHUNK[0] = chunk
  gp = gp_core.GP(w['X'], w['Y'], kernel.kernel_from_spec(w['kernel']), gp_core.ConstantMean(w['mean_const']), w['noise_var'])
  gp._fused_score(acq, cd[:200000]); torch.cuda.synchronize()
  ts = []
  for _ in range(3):
    t0 = time.perf_counter(); r = gp._fused_score(acq, cd); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
  print('chunk', int(gp._post.query('chunk')), 'ms', [round(t, 1) for t in ts], 'argmax', r[1], 'group', gp._post.query('last_c2_group'), flush=True)
  del gp
  device.release_workspaces()
