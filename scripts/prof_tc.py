"""Profiling driver: a few launches of the two big tensor-core convolutions at KITTI sizes (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops

dev = "cuda:0"
passes = int(os.environ.get("PASSES", "3"))
which = os.environ.get("WHICH", "both")
torch.manual_seed(0)
if which in ("both", "2d"):
  x2 = torch.randn(1, 376, 1248, 32, device=dev)
  w2 = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, device=dev) * 0.1)
  b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
  g2 = ops.geom((1, 376, 1248, 32), 3, dil=int(os.environ.get("DIL", "1")))
  for _ in range(3):
    ops.conv_c32_tc(x2, w2, g2, bias=b, scale=sc, shift=sh, residual=x2, lrelu=True, passes=passes)
if which in ("both", "3d"):
  x3 = torch.randn(1, 24, 47, 156, 32, device=dev)
  w3 = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05)
  b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
  g3 = ops.geom((1, 24, 47, 156, 32), 3)
  for _ in range(3):
    ops.conv_c32_tc(x3, w3, g3, bias=b, scale=sc, shift=sh, lrelu=True, passes=passes)
torch.cuda.synchronize()
print("ok")
