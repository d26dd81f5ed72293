"""torchrun check of the multi-GPU pieces on real GPUs (NCCL): sharded arg-max == single-rank arg-max,
sharded hp grid == local hp grid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dragonfly_b200 import synth_data, kernel, gp_core, device, hp_grid
from dragonfly_b200 import dist as D
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
w = synth_data.make_workload('c2_hartmann6_matern_ucb', n_train=1500, n_cand=40001)
k = w['kernel']
gp = gp_core.GP(w['X'], w['Y'], kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths']),
                gp_core.ConstantMean(w['mean_const']), w['noise_var'], device=local)
acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
C = w['candidates']
def score(lo, hi):
  b, i, _ = gp._fused_score(acq, C[lo:hi])
  return b, i
s, i = D.sharded_score_argmax(score, len(C), device=dev)
bs, bi, _ = gp._fused_score(acq, C)
assert i == bi and s == bs, (rank, i, bi, s, bs)
layout = hp_grid.EuclideanHPLayout(6, 'matern', nu=2.5)
rs = np.random.RandomState(0)
hps = np.concatenate((np.log(w['Y'].var()) + rs.uniform(-6, -3, (7, 1)), np.log(w['Y'].var()) + rs.uniform(-1, 1, (7, 1)),
                      rs.uniform(np.log(0.15), np.log(1.0), (7, 6))), axis=1)
lm, probs = hp_grid.sharded_lml_grid(w['X'], w['Y'], hps, layout, device=local)
lm0, _ = hp_grid.lml_for_hyperparams(w['X'], w['Y'], hps, layout, device=local)
assert (lm == lm0).all() and abs(probs.sum() - 1) < 1e-12
dist.barrier()
if rank == 0:
  print('DIST_OK world=%d argmax=%d' % (world, i))
dist.destroy_process_group()
