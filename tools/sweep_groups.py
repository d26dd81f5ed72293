"""Sweep the L2 scheduling group size of the two scoring kernels (N=5000, 2 chunks)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, device
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=5000, n_cand=6528 * 3)
k = w['kernel']
desc = kernel.build_descriptor(kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']))
cd = torch.from_numpy(w['candidates']).cuda()
for impl, opt, groups in [(1, 'i8_cb_group', [1, 2, 4, 6, 8, 12, 17, 34, 102]), (0, 'tma_cb_group', [1, 2, 3, 4, 6, 51])]:
  post = device.DevicePosterior(5000)
  post.set_option('score_impl', impl)
  post.set_kernel(desc); post.set_train(w['X'], w['Y'] - w['mean_const'])
  assert post.build(w['noise_var'])[0] == 0
  post.profile_enable(True)
  ref = None
  for g in groups:
    post.set_option(opt, g)
    post.eval(cd, mean_const=w['mean_const']); post.profile_read(1)
    mu, sd = post.eval(cd, mean_const=w['mean_const'])
    ms, nl, _ = post.profile_read(1)
    if ref is None: ref = sd.clone()
    print('impl %d %s=%3d: gemm-class %.3f ms/chunk  identical=%s' % (impl, opt, g, ms / nl, bool((sd == ref).all())))
