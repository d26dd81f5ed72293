"""
bench.py -- the headline metric of BASELINE.json: posterior + acquisition candidates / second at
N = 5000 training points, fp64 (SURVEY.md 8d "headline": Hartmann-6, Matern-2.5, EI).

One "step" = one BO inner-loop iteration over one batch of synthetic candidates:
    GP.build_posterior (K, Cholesky, L^-1, alpha, LML)  +  K_* rows + mu  +  |L^-1 k_*|^2 -> sigma
    +  EI  +  running arg-max                         for M candidates per GPU.
  value : candidates / s with the candidate matrix already resident in HBM (device tensor).
  e2e   : the same step through the public plugin call with HOST buffers (pinned NumPy candidates,
          training data uploaded, 16-byte result read back), copies inside the timed region.
Multi-GPU (torchrun, one rank per GPU): every rank builds the (tiny, replicated) posterior and
scores its own M candidates -- weak scaling, no data-path collective; the single collective is the
16-byte (score, index) all-gather of the final arg-max (dragonfly_b200/dist.py).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/gp_oracle.py, the
faithful chunked gp.eval(.., 'std') driver of SURVEY 8d) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'posterior+acq candidates/sec at N=5000 fp64'
UNIT = 'candidates/s'
CPU_CHUNK = 2000          # the reference's eval('std') builds chunk x chunk covariances


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=5)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--n-train', type=int, default=5000)
  p.add_argument('--cands-per-gpu', type=int, default=1000000)
  p.add_argument('--cpu-sample', type=int, default=24000)
  p.add_argument('--no-cpu-baseline', action='store_true')
  return p.parse_args()


def dist_env():
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  return rank, world, local


# ---------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference algorithm)
# ---------------------------------------------------------------------------------------------------
def use_all_host_threads():
  """ torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to run on all the host
      cores, so lift the BLAS / OpenMP pools back to the core count (threadpoolctl works after import). """
  try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=os.cpu_count())
  except Exception:  # pylint: disable=broad-except
    pass


def cpu_threads():
  try:
    from threadpoolctl import threadpool_info
    infos = [i for i in threadpool_info() if i.get('user_api') == 'blas']
    if infos:
      return max(i['num_threads'] for i in infos)
  except Exception:  # pylint: disable=broad-except
    pass
  return os.cpu_count()


def build_oracle_gp(w):
  from oracle import gp_oracle as O
  k = w['kernel']
  kern = O.OMaternKernel(k['dim'], k['nu'], k['scale'], k['dim_bandwidths'])
  mean_const = w['mean_const']
  t0 = time.perf_counter()
  gp = O.OGP(w['X'], w['Y'], kern, lambda x: np.array([mean_const] * len(x)), w['noise_var'])
  return O, gp, time.perf_counter() - t0


def cpu_score_sample(O, gp, cands, curr_best):
  t0 = time.perf_counter()
  val, idx, _ = O.chunked_scores(gp, cands, 'ei', chunk=CPU_CHUNK, curr_best=curr_best)
  return time.perf_counter() - t0, idx


def run_reference(args):
  rank, world, _ = dist_env()
  if rank != 0:
    return
  from dragonfly_b200 import synth_data
  use_all_host_threads()
  w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=args.n_train,
                               n_cand=args.cpu_sample)
  O, gp, build_s = build_oracle_gp(w)
  best = float(w['Y'].max())
  for _ in range(args.warmup):
    cpu_score_sample(O, gp, w['candidates'][:CPU_CHUNK], best)
  total = 0.0
  for _ in range(args.steps):
    dt, _ = cpu_score_sample(O, gp, w['candidates'], best)
    total += dt
  value = args.cpu_sample * args.steps / total
  cores = cpu_threads()
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
    'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
    'data': 'synthetic',
    'config': workload_config(args, 1),
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                     'sample': '%d candidates per step in chunks of %d through the faithful '
                               'gp.eval(chunk, "std") restatement (chunk x chunk covariance + TRSM), '
                               'EI + running arg-max; posterior build %.2f s excluded' % (
                                   args.cpu_sample, CPU_CHUNK, build_s)},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line))


def workload_config(args, world):
  return {'workload': 'headline: Hartmann-6 (d=6) Matern-2.5, N=%d train, EI over %d candidates per GPU '
                      '(BASELINE.json metric N=5000; configs[1] geometry at the metric\'s N)' % (
                          args.n_train, args.cands_per_gpu),
          'n_train': args.n_train, 'dim': 6, 'kernel': 'matern-2.5', 'acquisition': 'ei',
          'candidates_per_gpu': args.cands_per_gpu, 'global_candidates': args.cands_per_gpu * world,
          'parallelism': 'candidate-sharded x%d (posterior replicated)' % world,
          'step': 'build_posterior + score + arg-max',
          'l2': 'explicit L2 flush (512 MB write) before every timed step; per-step working set '
                '(W 210 MB + K_* chunk 267 MB) also exceeds the 126 MB L2'}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.path = '/tmp/dfb_clocks_%d_%d.csv' % (os.getpid(), gpu_index)
    self.proc = None
    self.gpu_index = gpu_index

  def start(self):
    try:
      self.f = open(self.path, 'w')
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '200'],
                                   stdout=self.f, stderr=subprocess.DEVNULL)
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def stop(self):
    out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
    if self.proc is None:
      return out
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    self.f.close()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for row in open(self.path):
      parts = [p.strip() for p in row.split(',')]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1])); mx.append(float(parts[2]))
      except ValueError:
        continue
      for name, val in zip(names, parts[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if sm:
      out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                 samples=len(sm))
    try:
      os.remove(self.path)
    except OSError:
      pass
    return out


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def measure_dgemm_peak(torch, dev):
  """ Live roofline denominator for the fp64 contraction: cuBLAS DGEMM 8192^3, best of 5
      (MEASURED_PEAKS.json holds no fp64 figure).  Not part of the product path. """
  n = 8192
  a = torch.randn(n, n, dtype=torch.float64, device=dev)
  b = torch.randn(n, n, dtype=torch.float64, device=dev)
  for _ in range(2):
    c = a @ b
  torch.cuda.synchronize(dev)
  best = 1e30
  for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(dev)
    best = min(best, e0.elapsed_time(e1))
  del a, b, c
  torch.cuda.empty_cache()
  return 2.0 * n ** 3 / best * 1e-9


def kstar_build_line(prof, n_train, peaks):
  """ The standalone materialising K_* build (fp64 rows written to HBM) against the HBM roofline. """
  ms, launches, cands = prof
  npad = (n_train + 127) // 128 * 128
  if ms <= 0:
    return None
  gbs = cands * npad * 8.0 / (ms * 1e-3) * 1e-9
  hbm = float(peaks.get('hbm_gbs', 6564.2))
  entries = cands * npad / (ms * 1e-3)
  return {'kernel': 'kstar_fast_kernel<Matern, p=2, d=6> writing fp64 K_* rows', 'achieved_gbs': gbs,
          'hbm_peak_gbs': hbm, 'frac_of_hbm': gbs / hbm, 'candidates_per_s': cands / (ms * 1e-3),
          'entries_per_s': entries, 'launch_ms_avg': ms / max(launches, 1),
          'bound': 'fp64 pipe / issue slots (sqrt + exp per entry), not HBM: see DESIGN.md 5.1'}


def run_ours(args):
  import torch
  import torch.distributed as dist
  from dragonfly_b200 import synth_data, kernel, gp_core, device
  from dragonfly_b200 import dist as dfb_dist
  rank, world, local = dist_env()
  assert torch.cuda.is_available(), 'bench.py needs a CUDA device: there is no CPU fallback'
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  # stdout carries exactly one JSON line: native libraries that write to fd 1 (NCCL prints its version banner
  # there) are pointed at stderr for the duration; the line itself goes to the saved descriptor.
  sys.stdout.flush()
  json_fd = os.dup(1)
  os.dup2(2, 1)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  M = args.cands_per_gpu
  w = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=args.n_train, n_cand=16)
  k = w['kernel']
  kern = kernel.MaternKernel(6, 2.5, k['scale'], k['dim_bandwidths'])
  mean = gp_core.ConstantMean(w['mean_const'])
  best_y = float(w['Y'].max())
  acq = device.make_acq_desc('ei', best=best_y)
  # this rank's candidate shard: global rows [rank*M, (rank+1)*M) of one seeded stream
  rs = np.random.RandomState(1000 + rank)
  cands_host = torch.empty((M, 6), dtype=torch.float64, pin_memory=True)
  cands_host.numpy()[:] = rs.random_sample((M, 6))
  cands_dev = cands_host.to(dev)
  flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
  Xh, Yh = np.ascontiguousarray(w['X']), np.ascontiguousarray(w['Y'])

  launches = [0]

  def one_step(cands):
    """ The public call sequence of one BO iteration (host X/Y in, arg-max point index out). """
    gp = gp_core.GP(Xh, Yh, kern, mean, w['noise_var'], device=local)
    best, idx, _ = gp._fused_score(acq, cands)
    gbest, gidx = dfb_dist.all_reduce_argmax(best, idx + rank * M, dev) if world > 1 else (best, idx)
    launches[0] += gp._post.launch_count()
    return gp, gbest, gidx

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def timed(cands, steps, profile=False):
    total_ms, prof = 0.0, None
    for _ in range(steps):
      flush.fill_(1.0)
      e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize(dev)
      t0 = time.perf_counter()
      e0.record()
      gp, gbest, gidx = one_step(cands)
      e1.record()
      torch.cuda.synchronize(dev)
      wall_ms = 1e3 * (time.perf_counter() - t0)
      # device events bracket the step; the host-side wall time is the same region seen from the
      # CPU (includes Python + ctypes overhead) -- report the larger so nothing is hidden.
      total_ms += max(e0.elapsed_time(e1), wall_ms)
      del gp
    return total_ms

  for _ in range(args.warmup):
    one_step(cands_dev)
  barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  launches[0] = 0
  ms_dev = timed(cands_dev, args.steps)
  n_launch = launches[0]
  barrier()
  ms_e2e = timed(cands_host.numpy(), args.steps)
  barrier()
  clocks = sampler.stop() if rank == 0 else None

  # per-kernel device timing for the roofline: same step, event pairs around every launch of the
  # dominant kernel on the launching stream (libdfb200's own profiling hooks).
  gp = gp_core.GP(Xh, Yh, kern, mean, w['noise_var'], device=local)
  gp._post.profile_enable(True)
  prof = {}
  for _ in range(2):
    flush.fill_(1.0)
    gp._fused_score(acq, cands_dev)
  for name, cls in [('kstar', 0), ('gemm', 1), ('acq', 2)]:
    prof[name] = gp._post.profile_read(cls)
  gp._post.profile_enable(False)
  used_i8 = bool(gp._post.query('last_used_i8'))
  shortlist = int(gp._post.query('last_shortlist'))
  i8_bound = gp._post.query('i8_sigma2_bound')
  i8_impl = int(gp._post.query('i8_impl'))
  i8_r256 = int(gp._post.query('i8_radix256'))
  del gp
  # the same step with the int8 path disabled: pure fp64 DMMA contraction, for reference
  device.DEFAULT_OPTIONS['score_impl'] = 0
  one_step(cands_dev)
  barrier()
  ms_fp64 = timed(cands_dev, 2)
  # the materialising K_* build of the fp64 path (the north-star's "K_* build vs HBM" figure), timed per
  # launch with the same event hooks: 8 * npad bytes written per candidate
  gp = gp_core.GP(Xh, Yh, kern, mean, w['noise_var'], device=local)
  gp._post.profile_enable(True)
  for _ in range(2):
    flush.fill_(1.0)
    gp._fused_score(acq, cands_dev[:200000])
  kstar64 = gp._post.profile_read(0)
  gp._post.profile_enable(False)
  del gp
  device.DEFAULT_OPTIONS.pop('score_impl')

  # incremental posterior update (dfb_extend_posterior) against the full rebuild the reference does on every
  # new observation: N-1 -> N points
  upd = {}
  if rank == 0:
    gp = gp_core.GP(Xh[:-1], Yh[:-1], kern, mean, w['noise_var'], device=local)
    post0 = gp._post
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    gp.add_data_multiple([Xh[-1]], [Yh[-1]])
    torch.cuda.synchronize(dev)
    upd['extend_1_point_ms'] = 1e3 * (time.perf_counter() - t0)
    upd['in_place'] = bool(gp._post is post0)
    lml_ext = gp.compute_log_marginal_likelihood()
    del gp, post0
    t0 = time.perf_counter()
    gp = gp_core.GP(Xh, Yh, kern, mean, w['noise_var'], device=local)
    torch.cuda.synchronize(dev)
    upd['full_build_ms'] = 1e3 * (time.perf_counter() - t0)
    upd['lml_rel_diff'] = abs(lml_ext - gp.compute_log_marginal_likelihood()) / abs(gp.compute_log_marginal_likelihood())
    upd['note'] = ('GP.add_data_multiple of one observation at N-1 -> N: in-place extension of the factorisation '
                   '(last row block of L / L^-1 only) vs GP(...) from scratch, host wall-clock incl. uploads')
    del gp

  t = torch.tensor([ms_dev, ms_e2e, ms_fp64], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_dev, ms_e2e, ms_fp64 = float(t[0]), float(t[1]), float(t[2])
  value = M * world * args.steps / (ms_dev * 1e-3)
  e2e_value = M * world * args.steps / (ms_e2e * 1e-3)

  if rank == 0:
    N = args.n_train
    gemm_ms, gemm_launches, gemm_cands = prof['gemm']
    flops_per_cand = float(N) * float(N + 1)       # triangular W: sum_i 2(i+1) = N(N+1) flops
    fp64_equiv = gemm_cands * flops_per_cand / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    dgemm_peak = measure_dgemm_peak(torch, dev)
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # pylint: disable=broad-except
      pass
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'gemm_traffic.json')
    if os.path.exists(tpath):
      try:
        traffic = json.load(open(tpath)).get({2: 'dram_bytes_per_launch_i8c2', 1: 'dram_bytes_per_launch_i8x2'}.get(i8_impl, 'dram_bytes_per_launch_i8')
                                             if used_i8 else 'dram_bytes_per_launch')
      except Exception:  # pylint: disable=broad-except
        traffic = None
    share = {kname: prof[kname][0] / max(sum(p[0] for p in prof.values()), 1e-9) for kname in prof}
    if used_i8:
      # dominant kernel: the tcgen05 int8 contraction (the digit planes of K_* are emitted by the K_*
      # kernel, those of W once per build).  Algorithmic work: one int8 digit product per kept (s, t) pair
      # for every fp64 multiply-add of the triangular contraction -- 15 with five radix-256 digits, 21 with
      # six radix-128 digits.
      n_products = 15.0 if i8_r256 else 21.0
      ops_per_cand = n_products * flops_per_cand
      achieved = gemm_cands * ops_per_cand / (gemm_ms * 1e-3) * 1e-12
      # int8 tensor peak: MEASURED_PEAKS.json has no int8 figure, so the denominator is the measured
      # tcgen05.mma kind::i8 issue rate of tools/ubench_i8.cu on this pool's B200
      # (profiles/r01_ubench_tcgen05_i8.txt): 4577 TOP/s for N >= 128 (2777 TOP/s for N = 64).
      peak, peak_n64 = 4577.2, 2777.1
      digits = ('five radix-256 digits, 15 exact int8 products' if i8_r256
                else 'six radix-128 digits, 21 exact int8 products')
      if i8_impl == 2:
        kname = ('score_i8c2_kernel (persistent 2-CTA clusters; tcgen05.mma.cta_group::2 kind::i8 M256 N128 K32 / '
                 'UTCIMMA.2CTA, four int32 accumulators = all 512 TMEM columns per SM, two passes per 256x128 '
                 'tile, 6-stage TMA ring): V = L^-1 K_*^T as %s, fused |v|^2' % digits)
      elif i8_impl == 1:
        kname = ('score_i8x2_kernel (persistent, one CTA per SM; tcgen05.mma kind::i8 M128 N128 K32 / UTCIMMA, four '
                 'int32 accumulators = all 512 TMEM columns, two passes per 128x128 tile, 4-stage TMA ring): '
                 'V = L^-1 K_*^T as %s, fused |v|^2' % digits)
      else:
        kname = ('score_i8_kernel (tcgen05.mma kind::i8 M128 N64 K32, six TMEM accumulators, TMA ring): '
                 'V = L^-1 K_*^T as %s, fused |v|^2' % digits)
      roofline = {
        'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TOP/s (int8)',
        'frac': achieved / peak, 'traffic': traffic, 'kernel': kname,
        'int8_products_per_fp64_fma': n_products, 'ops_per_candidate': ops_per_cand, 'launch_ms_avg': gemm_ms / max(gemm_launches, 1),
        'launches_timed': int(gemm_launches),
        'peak_source': 'measured tcgen05 kind::i8 issue rate, M128 N128/N256 K32, tools/ubench_i8.cu on this pool '
                       '(profiles/r01_ubench_tcgen05_i8.txt); MEASURED_PEAKS.json bf16 burst = %s TF/s for '
                       'comparison' % peaks.get('bf16_tflops', 'n/a'),
        'fp64_equivalent_tflops': fp64_equiv,
        'fp64_equivalent_vs_cublas_dgemm': fp64_equiv / dgemm_peak if dgemm_peak > 0 else None,
        'cublas_dgemm_tflops_live': dgemm_peak, 'share_of_scoring': share,
        'i8_sigma2_error_bound': i8_bound, 'argmax_shortlist_rescored_fp64': shortlist}
      if i8_impl == 0:
        roofline['frac_of_shape_limited_peak'] = achieved / peak_n64
    else:
      roofline = {
        'bound': 'tensor', 'achieved': fp64_equiv, 'peak': dgemm_peak, 'unit': 'TFLOP/s',
        'frac': fp64_equiv / dgemm_peak if dgemm_peak > 0 else None, 'traffic': traffic,
        'kernel': 'score_tma_kernel (fp64 DMMA, TMA + mbarrier ring: V = L^-1 K_*^T fused with |v|^2)',
        'flops_per_candidate': flops_per_cand, 'launch_ms_avg': gemm_ms / max(gemm_launches, 1),
        'launches_timed': int(gemm_launches),
        'peak_source': 'live cuBLAS DGEMM 8192^3 burst on this GPU (MEASURED_PEAKS.json has no fp64 '
                       'figure; DMMA issue peak measured 37.1 TFLOP/s)', 'share_of_scoring': share}
    step_ms = ms_dev / args.steps
    line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None,
      'dtype': 'f64' + (' (sigma^2 contraction: int8 digit expansion of the fp64 operands on tcgen05, a-priori '
                        '|d sigma^2| <= %.1e, measured 3.6e-10; arg-max re-scored in fp64 DMMA)' % i8_bound
                        if used_i8 else ''),
      'data': 'synthetic', 'config': workload_config(args, world),
      'clocks': clocks,
      'e2e': {'value': e2e_value, 'unit': UNIT,
              'h2d_bytes_per_step': int(M * 6 * 8 + N * 6 * 8 + N * 8),
              'd2h_bytes_per_step': 16},
      'gpu_launches': int(n_launch),
      'roofline': roofline,
      'fp64_dmma_only': {'value': M * world * 2 / (ms_fp64 * 1e-3), 'unit': UNIT,
                         'note': 'same step with DFB200_SCORE=fp64 (no int8 path), 2 timed steps'},
      'kstar_build_fp64': kstar_build_line(kstar64, N, peaks),
      'posterior_update': upd,
    }
    if not args.no_cpu_baseline:
      use_all_host_threads()
      wc = synth_data.make_workload('headline_hartmann6_matern_ei', n_train=args.n_train,
                                    n_cand=args.cpu_sample)
      O, ogp, build_s = build_oracle_gp(wc)
      dt, _ = cpu_score_sample(O, ogp, wc['candidates'], best_y)
      line['cpu_baseline'] = {
        'value': args.cpu_sample / dt, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
        'sample': '%d candidates in chunks of %d through the faithful gp.eval(chunk, "std") '
                  'restatement + EI + arg-max (%.1f s); posterior build %.2f s excluded' % (
                      args.cpu_sample, CPU_CHUNK, dt, build_s)}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + '\n').encode())
  os.close(json_fd)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
