"""Per-role cycle counters of the ws convolution (snb_conv_c32_ws_profile).  Launches are back to back with programmatic
dependent launch, so w_start / kernel_total include the wait for the previous launch: read mma_total / cv_total / epi_total.
Counters are means over 148 rows; scale by 148 / grid for grids smaller than 148.
(The round-2 knock-out experiments — converters / epilogue / TMA idled one at a time through a debug kernel parameter — used this
script with a temporary SNB200_WS_KNOCK switch in the kernel; the switch itself cost 1.4 % of the forward and was removed.  The
numbers are in DESIGN.md section 4.2.)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops, _cabi
from stereonet_b200._cabi import ConvEpilogue
dev = "cuda:0"
torch.manual_seed(0)
names = ["prod_wait_rempty", "prod_total", "mma_wait_afull", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_total", "epi_bar",
         "w_start", "w_wait", "epi_out", "cv_wait_rfull", "cv_wait_aempty", "cv_total", "cv_wait_res", "kernel_total"]
for shape in ((1, 376, 1248, 32), (1, 24, 47, 156, 32)):
  ks = (3, 3, 3) if len(shape) == 5 else (3, 3)
  x = torch.randn(shape, device=dev); w = torch.randn(32, 32, *ks, device=dev) * 0.05
  b = torch.randn(32, device=dev)
  y = torch.empty_like(x)
  g = ops.geom(shape, 3, dil=1)
  wimg = ops.prep_conv_weights_tc(w, fmt="ws")
  for knock in (0,):
    cnt = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    e = ConvEpilogue(ops._p(b), None, None, None, None, 1)
    for _ in range(3):
      _cabi.check(_cabi.lib().snb_conv_c32_ws_profile(ops._p(x), ops._p(wimg), ops._p(y), C.byref(g), C.byref(e), ops._p(cnt), ops._stream(x)), "profile")
    torch.cuda.synchronize()
    m = cnt.double().mean(0).tolist()
    print(f"{shape} knock={knock}: " + "  ".join(f"{n}={v/1e3:.1f}k" for n, v in zip(names, m)), flush=True)
