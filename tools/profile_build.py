"""Posterior build only (N=5000), for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, device
w = synth_data.make_workload('headline_hartmann6_matern_ei', n_cand=16)
k = w['kernel']
post = device.DevicePosterior(5000)
post.set_option('score_impl', 0)
post.set_kernel(kernel.build_descriptor(kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths'])))
post.set_train(w['X'], w['Y'] - w['mean_const'])
for _ in range(2):
  info, lml = post.build(w['noise_var'], 0.0, 0)
print('ok', info, lml)
