"""ncu driver: the hot kernels of the path at KITTI size, a few launches each (forward convs, wgrad, cost volume, head)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
x2 = torch.randn(1, 376, 1248, 32, device=dev); d2 = torch.randn(1, 376, 1248, 32, device=dev)
w2 = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, device=dev) * 0.1)
x3 = torch.randn(1, 24, 47, 156, 32, device=dev); d3 = torch.randn(1, 24, 47, 156, 32, device=dev)
w3 = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05)
wa = torch.randn(1, 32, 3, 3, 3, device=dev)
fl = torch.randn(1, 47, 156, 32, device=dev); fr = torch.randn(1, 47, 156, 32, device=dev)
fl8 = torch.randn(8, 47, 156, 32, device=dev); fr8 = torch.randn(8, 47, 156, 32, device=dev)
g2 = ops.geom((1, 376, 1248, 32), 3, dil=1); g28 = ops.geom((1, 376, 1248, 32), 3, dil=8); g3 = ops.geom((1, 24, 47, 156, 32), 3)
img = torch.rand(2, 3, 376, 1248, device=dev); w5 = torch.randn(32, 3, 5, 5, device=dev) * 0.2
coarse = torch.rand(1, 47, 156, device=dev) * 20; rgb = img[:1].contiguous(); w4 = torch.randn(32, 4, 3, 3, device=dev) * 0.2
for rep in range(3):
  flush.zero_(); ops.conv_c32_tc(x2, w2, g2, bias=b, scale=sc, shift=sh, residual=x2, lrelu=True)
  flush.zero_(); ops.conv_c32_tc(x2, w2, g28, bias=b, scale=sc, shift=sh, residual=x2, lrelu=True)
  flush.zero_(); ops.conv_c32_tc(x3, w3, g3, bias=b, scale=sc, shift=sh, lrelu=True)
  flush.zero_(); ops.conv_c32_wgrad_tc(x2, d2, g2, (32, 32, 3, 3))
  flush.zero_(); ops.conv_c32_wgrad_tc(x3, d3, g3, (32, 32, 3, 3, 3))
  flush.zero_(); ops.cost_volume(fl, fr, 24)
  flush.zero_(); ops.cost_volume(fl8, fr8, 24)
  flush.zero_(); taps = ops.conv_c32_taps(x3, wa, 27)
  flush.zero_(); ops.tapsum_softargmin(taps, b[:1].contiguous(), True)
  flush.zero_(); ops.conv5x5s2_c3_phases(img, w5, b)
  flush.zero_(); ops.refine_in_conv(coarse, rgb, w4, b, scale=sc, shift=sh, lrelu=True)
torch.cuda.synchronize()
print("ok")
