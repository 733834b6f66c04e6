"""Per-role cycle counters of the flat-tiled loader-warp conv kernel on a 2-D problem (diagnostics; the 2-D walk kernel:
tc2d_counters.py, the 3-D TMA kernel: tc3_counters.py)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops, _cabi
from stereonet_b200._cabi import ConvEpilogue
dev = "cuda:0"
torch.manual_seed(0)
names = ["ld_wait_empty", "ld_total", "mma_wait_full", "mma_wait_tmem", "mma_total", "epi_wait_tfull", "epi_total", "epi_bar", "epi_pre", "epi_tmem2smem", "epi_out", "", "", "", "", ""]
def run(tag, x, w, g, residual, passes):
  wimg = ops.prep_conv_weights_tc(w)
  b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
  y = torch.empty_like(x)
  cnt = torch.zeros(148, 16, dtype=torch.int64, device=dev)
  e = ConvEpilogue(ops._p(b), ops._p(sc), ops._p(sh), ops._p(x if residual else None), None, 1)
  for _ in range(3):
    _cabi.check(_cabi.lib().snb_conv_c32_tc_profile(ops._p(x), ops._p(wimg), ops._p(y), C.byref(g), C.byref(e), passes, ops._p(cnt), ops._stream(x)), "profile")
  torch.cuda.synchronize()
  nt = _cabi.lib().snb_conv_c32_tc_num_tiles(C.byref(g))
  m = cnt.double().mean(0).tolist()
  print(f"{tag} passes={passes} tiles={nt} per-SM tiles={nt/148:.1f}: " + "  ".join(f"{n}={v/1e3:.1f}k" for n, v in zip(names, m) if n))
for passes in (3, 1):
  x2 = torch.randn(1, 376, 1248, 32, device=dev)
  run("2d dil1", x2, torch.randn(32, 32, 3, 3, device=dev) * 0.1, ops.geom((1, 376, 1248, 32), 3, dil=1), True, passes)
