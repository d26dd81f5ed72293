"""
CPU-side checks of the drop-in boundary: libdfb200.so loads, exports every symbol that
include/dfb200.h declares, the ctypes prototype table covers exactly those symbols, the POD structs
have the C layout, and -- with no GPU in this container -- handle creation fails loudly instead of
falling back to anything.  No compute calls.
"""
import ctypes as C
import os
import re

import pytest

from dragonfly_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'dfb200.h')


def header_functions():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(dfb_[a-z_0-9]+)\s*\(', src)))


def test_library_is_built_in_tree():
  assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
  assert os.path.dirname(_lib.LIB_PATH).endswith('dragonfly_b200')


def test_every_declared_symbol_is_exported():
  lib = C.CDLL(_lib.LIB_PATH)
  names = header_functions()
  assert len(names) >= 20
  for name in names:
    assert hasattr(lib, name), 'libdfb200.so does not export %s' % (name)


def test_prototype_table_matches_header():
  assert sorted(_lib.PROTOTYPES.keys()) == header_functions()
  _lib.load()


def test_struct_layouts_match_the_header_constants():
  src = open(HEADER).read()
  consts = dict((k, int(v)) for k, v in re.findall(r'#define\s+(DFB_[A-Z_0-9]+)\s+(\d+)\b', src))
  for name in ['DFB_MAX_FACTORS', 'DFB_MAX_TERMS', 'DFB_MAX_SLOTS', 'DFB_MAX_MATERN_P', 'DFB_HOST',
               'DFB_DEVICE', 'DFB_ACQ_UCB', 'DFB_ACQ_EI', 'DFB_ACQ_PI', 'DFB_ACQ_TTEI',
               'DFB_BUILD_FULL', 'DFB_BUILD_LML_ONLY', 'DFB_BUILD_NO_ALPHA', 'DFB_BASE_SE',
               'DFB_BASE_MATERN']:
    assert getattr(_lib, name) == consts[name], name
  assert C.sizeof(_lib.FactorDesc) == 16 + 8 * 4 + 8 * (consts['DFB_MAX_MATERN_P'] + 1)
  expect = 24 + 16 + 4 * (consts['DFB_MAX_TERMS'] + 1)
  expect = (expect + 7) // 8 * 8
  expect += 8 * consts['DFB_MAX_TERMS'] + C.sizeof(_lib.FactorDesc) * consts['DFB_MAX_FACTORS']
  expect += (4 + 4 + 8) * consts['DFB_MAX_SLOTS']
  assert C.sizeof(_lib.KernelDesc) == expect
  assert C.sizeof(_lib.AcqDesc) == 8 + 4 * 8


def test_version_and_error_string():
  lib = _lib.load()
  assert lib.dfb_version() == 100
  assert isinstance(_lib.last_error(), str)


def test_no_gpu_means_loud_failure_not_fallback():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is visible here')
  lib = _lib.load()
  hp = C.c_void_p()
  status = lib.dfb_create(C.byref(hp), 0)
  assert status < 0 and not hp.value
  assert 'no CPU fallback' in _lib.last_error()
  from dragonfly_b200 import device
  with pytest.raises(RuntimeError):
    device.DevicePosterior(16)


def test_missing_library_raises_importerror(monkeypatch):
  monkeypatch.setattr(_lib, '_lib', None)
  monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libdfb200.so')
  with pytest.raises(ImportError):
    _lib.load()
