"""Per-class device times of the scoring chunk loop at the headline size (int8 path), for A/B runs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, device

def main():
  n_train, n_cand = 5000, 6528 * int(os.environ.get('CHUNKS', '8'))
  w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=n_train, n_cand=n_cand)
  k = w['kernel']
  desc = kernel.build_descriptor(kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']))
  cd = torch.from_numpy(w['candidates']).cuda()
  post = device.DevicePosterior(n_train)
  post.set_option('score_impl', 1)
  post.set_option('i8_impl', int(os.environ.get('I8_IMPL', '1')))
  if 'CB_GROUP' in os.environ: post.set_option('i8_cb_group', int(os.environ['CB_GROUP']))
  post.set_kernel(desc)
  post.set_train(w['X'], w['Y'] - w['mean_const'])
  info, lml = post.build(w['noise_var'])
  assert info == 0
  acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
  for rep in range(3):
    post.profile_enable(True)
    post.score_argmax(acq, cd, mean_const=w['mean_const'])
    torch.cuda.synchronize()
    r = [post.profile_read(c) for c in range(3)]
  print('impl %s dbg %s: per chunk ms: kstar %.3f gemm %.3f acq %.3f' % (
      os.environ.get('I8_IMPL', '1'), os.environ.get('DFB200_I8_DBG', '0'),
      r[0][0] / r[0][1], r[1][0] / r[1][1], r[2][0] / max(r[2][1], 1)))

if __name__ == '__main__':
  main()
