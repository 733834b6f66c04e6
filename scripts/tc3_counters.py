"""Per-role wait counters of the TMA 3-D conv kernel (diagnostics)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops, _cabi
from stereonet_b200._cabi import ConvEpilogue
dev = "cuda:0"
torch.manual_seed(0)
names = ["prod_wait_rempty", "prod_wait_bempty", "prod_total", "mma_wait_afull", "mma_wait_bfull", "mma_wait_tempty", "mma_total",
         "cv_wait_rfull", "cv_wait_aempty", "cv_wait_st", "cv_total", "epi_wait_tfull", "epi_total"]
x = torch.randn(1, 24, 47, 156, 32, device=dev)
wimg = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05)
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
y = torch.empty_like(x)
g = ops.geom(tuple(x.shape), 3)
for passes in (3, 1):
  cnt = torch.zeros(148, 16, dtype=torch.int64, device=dev)
  e = ConvEpilogue(ops._p(b), ops._p(sc), ops._p(sh), None, None, 1)
  for _ in range(3):
    _cabi.check(_cabi.lib().snb_conv_c32_tc_profile(ops._p(x), ops._p(wimg), ops._p(y), C.byref(g), C.byref(e), passes, ops._p(cnt), ops._stream(x)), "profile")
  torch.cuda.synchronize()
  m = cnt.double().mean(0).tolist()
  print(f"passes={passes}: " + "  ".join(f"{n}={v/1e3:.1f}k" for n, v in zip(names, m)))
