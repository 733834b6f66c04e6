import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
for (B, H, W, dil) in [(2, 17, 29, 4), (2, 17, 29, 1), (1, 17, 29, 4), (2, 17, 29, 8), (2, 40, 60, 4)]:
  x = torch.randn(B, H, W, 32, device=dev)
  r = torch.randn(B, H, W, 32, device=dev)
  w = torch.randn(32, 32, 3, 3, device=dev) * 0.1
  g = ops.geom((B, H, W, 32), 3, dil=dil)
  for mode in (0, 1):
    ref, _ = ops.conv_c32(x, ops.prep_conv_weights(w, mode), g, residual=r)
    ref0, _ = ops.conv_c32(x, ops.prep_conv_weights(w, mode), g)
    for it in range(3):
      y, _ = ops.conv_c32_tc(x, ops.prep_conv_weights_tc(w, mode), g, residual=r)
      y0, _ = ops.conv_c32_tc(x, ops.prep_conv_weights_tc(w, mode), g)
      d = (y - ref).abs(); d0 = (y0 - ref0).abs()
      idx = torch.nonzero(d > 1e-3)
      print(f"B{B} H{H} W{W} dil{dil} mode{mode} it{it}: res diff {d.max().item():.3e}  nores diff {d0.max().item():.3e}  nbad {idx.shape[0]}",
            idx[:4].tolist() if idx.shape[0] else "")
