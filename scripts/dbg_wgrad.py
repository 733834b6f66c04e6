"""Per-tap error pattern of the tensor-core wgrad kernel (diagnostics for the MN-major / shifted-descriptor scheme)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from stereonet_b200 import ops
dev = "cuda:0"
def cl(x):
  perm = (0, 2, 3, 1) if x.dim() == 4 else (0, 2, 3, 4, 1)
  return x.permute(*perm).contiguous().to(dev)
def run(B, D, H, W, dil, passes):
  three_d = D > 1
  gen = torch.Generator().manual_seed(1)
  shp = (B, 32, D, H, W) if three_d else (B, 32, H, W)
  x = torch.randint(-3, 4, shp, generator=gen).float(); dz = torch.randint(-3, 4, shp, generator=gen).float()
  w = torch.zeros((32, 32, 3, 3, 3) if three_d else (32, 32, 3, 3), requires_grad=True)
  y = F.conv3d(x, w, padding=1) if three_d else F.conv2d(x, w, padding=dil, dilation=dil)
  y.backward(dz)
  ref = w.grad
  xg, dzg = cl(x), cl(dz)
  g = ops.geom(xg.shape, 3, stride=1, dil=dil)
  got = ops.conv_c32_wgrad_tc(xg, dzg, g, tuple(ref.shape), passes=passes).cpu()
  torch.cuda.synchronize()
  err = (got - ref).abs()
  per_tap = err.flatten(2).amax(dim=(0, 1))
  print(f"B{B} D{D} H{H} W{W} dil{dil} passes{passes}: max err {err.max().item():.3e} ref max {ref.abs().max().item():.3e}")
  print("   per-tap max err:", [f"{v:.2g}" for v in per_tap.tolist()])
for cfg in [(1, 1, 5, 16, 1, 1), (1, 1, 5, 16, 8, 1), (1, 1, 17, 29, 1, 1), (1, 1, 17, 29, 2, 3), (1, 1, 33, 141, 4, 3), (1, 6, 9, 21, 1, 1), (1, 6, 9, 21, 1, 3)]:
  try:
    run(*cfg)
  except Exception as e:
    print(cfg, "FAILED:", e)
    break
