"""3-D tensor-core conv: TMEM-A vs smem-A operand path, parity against the FFMA kernel and timing at KITTI size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
from stereonet_b200 import ops
from bench import time_kernel
dev = "cuda:0"
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
stream = torch.cuda.current_stream()
for shape in [(1, 3, 3, 8, 32), (1, 6, 9, 21, 32), (1, 24, 47, 156, 32), (2, 24, 47, 156, 32), (1, 24, 68, 120, 32)]:
  x = torch.randn(shape, device=dev)
  w = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
  b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
  g = ops.geom(shape, 3)
  wimg = ops.prep_conv_weights_tc(w); wp = ops.prep_conv_weights(w)
  ref, _ = ops.conv_c32(x, wp, g, bias=b, scale=sc, shift=sh, lrelu=True)
  flops = 2 * 27 * 32 * 32 * x.numel() / 32
  for passes in (3, 1):
    for a_smem in (False, True):
      y, st = ops.conv_c32_tc(x, wimg, g, bias=b, scale=sc, shift=sh, lrelu=True, passes=passes, legacy3d=a_smem)
      err = (y - ref).abs().max().item() / ref.abs().max().item()
      ms, med = time_kernel(lambda: ops.conv_c32_tc(x, wimg, g, bias=b, scale=sc, shift=sh, lrelu=True, passes=passes, legacy3d=a_smem), 10, flush, stream)
      print(f"{shape} passes{passes} {'legacy' if a_smem else 'tma   '}: {ms*1e3:7.1f} us  {flops/ms/1e9:6.1f} TFLOP/s  rel err {err:.2e}", flush=True)
