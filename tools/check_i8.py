"""Diagnostics for the int8-slice tcgen05 scoring path against the fp64 DMMA path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, device, _lib

def run(n_train, n_cand, chunk=0):
  w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=n_train, n_cand=n_cand)
  k = w['kernel']
  desc = kernel.build_descriptor(kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']))
  out = []
  cd = torch.from_numpy(w['candidates']).cuda()
  for impl in (0, 1):
    post = device.DevicePosterior(n_train, chunk=chunk)
    post.set_option('score_impl', impl)
    post.set_option('i8_ts', int(os.environ.get('I8_TS', '0')))
    post.set_option('i8_fuse', int(os.environ.get('I8_FUSE', '1')))
    post.set_option('i8_impl', int(os.environ.get('I8_IMPL', '0')))
    post.set_kernel(desc)
    post.set_train(w['X'], w['Y'] - w['mean_const'])
    info, lml = post.build(w['noise_var'])
    assert info == 0
    post.profile_enable(True)
    mu, sd = post.eval(cd, mean_const=w['mean_const'])
    torch.cuda.synchronize()
    ms, nl, units = post.profile_read(1)
    ms_k, _, _ = post.profile_read(0)
    ms += ms_k
    acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
    bs, bi, _ = post.score_argmax(acq, cd, mean_const=w['mean_const'])
    out.append((mu.cpu().numpy(), sd.cpu().numpy(), bs, bi, ms, nl))
  (mu0, sd0, bs0, bi0, ms0, nl0), (mu1, sd1, bs1, bi1, ms1, nl1) = out
  dv = np.abs(sd0 ** 2 - sd1 ** 2)
  print('N=%d M=%d scale=%.4f: mu equal %s | max|dvar| %.3e (rel to scale %.3e) mean %.3e | nan %d/%d | argmax %d vs %d | kstar+gemm ms fp64 %.3f (%d launches) i8 %.3f (%d)' % (
      n_train, n_cand, k['scale'], bool((mu0 == mu1).all()), np.nanmax(dv), np.nanmax(dv) / k['scale'], np.nanmean(dv),
      int(np.isnan(sd1).sum()), int(np.isnan(sd0).sum()), bi0, bi1, ms0, nl0, ms1, nl1))
  idx = int(np.nanargmax(dv))
  print('   worst idx %d: var fp64 %.12e i8 %.12e' % (idx, sd0[idx] ** 2, sd1[idx] ** 2))

if __name__ == '__main__':
  run(100, 256)
  run(300, 1000)
  run(700, 3000, chunk=1024)
  run(5000, 6528 * 2)
