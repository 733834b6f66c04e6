"""CPU-only: the C-ABI library loads and exports every symbol include/snb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  names = []
  for fn in os.listdir(os.path.join(ROOT, "include")):
    if fn.endswith(".h"):
      src = open(os.path.join(ROOT, "include", fn)).read()
      names += re.findall(r"SNB_API\s+[\w\s\*]+?\b(snb_\w+)\s*\(", src)
  return sorted(set(names))


def test_header_declares_something():
  assert len(declared_symbols()) >= 15


def test_library_exports_every_declared_symbol():
  import __graft_entry__
  __graft_entry__.build()
  from stereonet_b200 import _cabi
  lib = ctypes.CDLL(_cabi.LIB_PATH)
  missing = [n for n in declared_symbols() if not hasattr(lib, n)]
  assert not missing, missing


def test_binding_table_matches_header():
  from stereonet_b200 import _cabi
  assert sorted(_cabi.SIGNATURES) == declared_symbols()
  l = _cabi.lib()
  assert l.snb_version() >= 100
  assert isinstance(l.snb_last_error(), bytes)


def test_num_tiles_helpers_are_host_only():
  from stereonet_b200 import _cabi, ops
  g = ops.geom((1, 47, 156, 32), 3)
  assert _cabi.lib().snb_conv_c32_num_tiles(ctypes.byref(g)) == (47 * 156 + 127) // 128
  g3 = ops.geom((2, 24, 47, 156, 32), 3)
  assert (g3.OD, g3.OH, g3.OW, g3.KD) == (24, 47, 156, 3)
  g5 = ops.geom((1, 135, 240, 32), 5, stride=2, pad=2)
  assert (g5.OH, g5.OW) == (68, 120)           # 540 -> 270 -> 135 -> 68
