"""A/B of the pair kernel's L2 eviction-priority hints (option i8_l2_hint) at the headline geometry."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, gp_core, device
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=5000, n_cand=16)
gp = gp_core.GP(w['X'], w['Y'], kernel.kernel_from_spec(w['kernel']), gp_core.ConstantMean(w['mean_const']), w['noise_var'])
acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
cd = torch.from_numpy(np.random.RandomState(1000).random_sample((1000000, 6))).cuda()
for rep in range(2):
  for hint in (0, 1, 2):
    gp._post.set_option('i8_l2_hint', hint)
    gp._fused_score(acq, cd[:100000]); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
      t0 = time.perf_counter(); r = gp._fused_score(acq, cd); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    gp._post.profile_enable(True); gp._fused_score(acq, cd)
    g = gp._post.profile_read(1); k = gp._post.profile_read(0); gp._post.profile_enable(False)
    print('hint', hint, 'ms', [round(t, 1) for t in ts], 'gemm/launch %.4f' % (g[0] / g[1]), 'kstar/launch %.4f' % (k[0] / k[1]), 'argmax', r[1], flush=True)
