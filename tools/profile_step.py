"""One short pass of the hot path for ncu captures: build the N=5000 posterior once, then score a few
chunks of candidates with EI (device-resident).  Usage: python tools/profile_step.py [n_cand] [n_train]"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, gp_core, device

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 6528 * 3
n_train = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=n_train, n_cand=n_cand)
k = w['kernel']
gp = gp_core.GP(w['X'], w['Y'], kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                gp_core.ConstantMean(w['mean_const']), w['noise_var'], device=0)
acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
cd = torch.from_numpy(w['candidates']).cuda()
for _ in range(2):
  best, idx, _ = gp._fused_score(acq, cd)
torch.cuda.synchronize()
print('ok', best, idx, gp._post.launch_count())
