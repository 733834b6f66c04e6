"""ncu driver: one launch each of the 2-D conv with A from TMEM and A from smem (3 passes and 1 pass), KITTI size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(1, 376, 1248, 32, device=dev)
w = ops.prep_conv_weights_tc(torch.randn(32, 32, 3, 3, device=dev) * 0.1)
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
g = ops.geom((1, 376, 1248, 32), 3, dil=1)
for rep in range(2):
  for passes in (3, 1):
    for a_smem in (False, True):
      ops.conv_c32_tc(x, w, g, bias=b, scale=sc, shift=sh, residual=x, lrelu=True, passes=passes, a_smem=a_smem)
torch.cuda.synchronize()
print("ok")
