"""
CPU restatement (NumPy) of the counter-based generator behind dfb_fill_rng -- TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
as 1, 2, 3", SC'11; Random123 known-answer vectors in tests/test_oracle_golden.py), the 53-bit uniform map and
Box-Muller exactly as dragonfly_b200/csrc/kernels.cu does them:
    counter = (col & 0xffffffff, col >> 32, draw index s, what), key = (seed & 0xffffffff, seed >> 32)
    u1 = ((r0 << 32 | r1) >> 11 + 1/2) 2^-53,  u2 likewise from (r2, r3),  z = sqrt(-2 ln u1) cos(2 pi u2)
This is not part of the reference (which uses np.random.normal, general_utils.py:230): it pins OUR device generator.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """ Vectorised over uint32 arrays; returns four uint32 arrays. """
  c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3)]
  k0 = np.asarray(k0, dtype=np.uint32).copy(); k1 = np.asarray(k1, dtype=np.uint32).copy()
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = M0 * c0.astype(np.uint64)
      p1 = M1 * c2.astype(np.uint64)
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
      n0 = hi1 ^ c1 ^ k0
      n2 = hi0 ^ c3 ^ k1
      c0, c1, c2, c3 = n0, lo1, n2, lo0
      k0 = (k0 + W0).astype(np.uint32); k1 = (k1 + W1).astype(np.uint32)
  return c0, c1, c2, c3


def u53(hi, lo):
  v = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)
  return (v.astype(np.float64) + 0.5) * 2.0 ** -53


def fill(seed, col0, S, m, what=0):
  """ The S x m matrix dfb_fill_rng(seed, col0, S, m, what) writes: what = 0 normals, 1 uniforms. """
  s_idx, a_idx = np.meshgrid(np.arange(S, dtype=np.uint64), np.arange(m, dtype=np.uint64), indexing='ij')
  col = a_idx + np.uint64(col0)
  r = philox4x32_10((col & MASK).astype(np.uint32), (col >> np.uint64(32)).astype(np.uint32),
                    s_idx.astype(np.uint32), np.full(s_idx.shape, what, dtype=np.uint32),
                    np.full(s_idx.shape, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(s_idx.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32))
  u1 = u53(r[0], r[1])
  if what == 1:
    return u1
  u2 = u53(r[2], r[3])
  return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
