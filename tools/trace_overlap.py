"""Do the K_* CTAs run beside the persistent contraction CTAs?  Arms libdfb200's per-CTA trace (dfb_debug_trace), runs one
overlapped scoring call over 6 chunks and prints, per kernel launch, when its CTAs ran and on which SMs -- and for every
K_* CTA whether a contraction CTA was resident on the same SM at that time."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dragonfly_b200 import synth_data, kernel, gp_core, device, _lib

w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=5000, n_cand=16)
gp = gp_core.GP(w['X'], w['Y'], kernel.kernel_from_spec(w['kernel']), gp_core.ConstantMean(w['mean_const']), w['noise_var'])
acq = device.make_acq_desc('ei', best=float(w['Y'].max()))
cd = torch.from_numpy(np.random.RandomState(1000).random_sample((6 * 6528, 6))).cuda()
gp._post.set_option('kstar_overlap', 1)
gp._fused_score(acq, cd); torch.cuda.synchronize()
cap = 40000
buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device='cuda')
lib = _lib.load()
assert lib.dfb_debug_trace(C.c_void_p(buf.data_ptr()), cap) == 0
gp._fused_score(acq, cd); torch.cuda.synchronize()
assert lib.dfb_debug_trace(None, 0) == 0
b = buf.cpu().numpy()
n = min(int(b[0]), cap)
rec = b[1:1 + 4 * n].reshape(n, 4)
kind = rec[:, 0] >> 32; sm = rec[:, 0] & 0xffffffff
t0 = rec[:, 1] - rec[:, 1].min(); t1 = rec[:, 2] - rec[:, 1].min()
out = {'records': n}
g = np.where(kind == 2)[0]; k = np.where(kind == 1)[0]
out['contraction_ctas'] = len(g); out['kstar_ctas'] = len(k)
# group contraction CTAs into launches by start time gaps
order = g[np.argsort(t0[g])]
launches = []
for i in order:
  if not launches or t0[i] > launches[-1]['end'] - 1000 and t0[i] - launches[-1]['start'] > 200000:
    launches.append({'start': int(t0[i]), 'end': int(t1[i]), 'n': 1})
  else:
    launches[-1]['end'] = max(launches[-1]['end'], int(t1[i])); launches[-1]['n'] += 1
out['contraction_launches_us'] = [(l['start'] / 1e3, l['end'] / 1e3, l['n']) for l in launches]
# for each K_* CTA: was a contraction CTA resident on the same SM over its whole lifetime?
inside = 0
for i in k:
  same = g[sm[g] == sm[i]]
  if np.any((t0[same] <= t0[i]) & (t1[same] >= t1[i])):
    inside += 1
out['kstar_ctas_fully_inside_a_contraction_cta_on_the_same_sm'] = inside
out['kstar_cta_us'] = {'median': float(np.median(t1[k] - t0[k])) / 1e3, 'p90': float(np.percentile(t1[k] - t0[k], 90)) / 1e3}
out['kstar_span_us'] = (float(t0[k].min()) / 1e3, float(t1[k].max()) / 1e3)
hist, edges = np.histogram(t0[k] / 1e3, bins=24)
out['kstar_start_hist_us'] = [(round(float(edges[i])), int(hist[i])) for i in range(len(hist))]
print(json.dumps(out, indent=1))
json.dump(out, open('gpurun_out/trace_overlap.json', 'w'), indent=1)
