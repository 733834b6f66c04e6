"""Time the weight-gradient kernels at KITTI sizes (L2 flushed before every launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
from stereonet_b200 import ops
from bench import time_kernel
dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
stream = torch.cuda.current_stream()
for shape, dil in [((1, 376, 1248, 32), 1), ((1, 376, 1248, 32), 2), ((1, 376, 1248, 32), 4), ((1, 376, 1248, 32), 8), ((1, 24, 47, 156, 32), 1), ((1, 47, 156, 32), 1)]:
  x = torch.randn(shape, device=dev); dz = torch.randn(shape, device=dev)
  g = ops.geom(shape, 3, stride=1, dil=dil)
  ws = (32, 32, 3, 3, 3) if len(shape) == 5 else (32, 32, 3, 3)
  taps = 27 if len(shape) == 5 else 9
  flops = 2 * taps * 32 * 32 * x.numel() / 32
  for name, fn in [("tc3", lambda: ops.conv_c32_wgrad_tc(x, dz, g, ws, 3)), ("tc1", lambda: ops.conv_c32_wgrad_tc(x, dz, g, ws, 1))] + ([("ffma", lambda: ops.conv_c32_wgrad(x, dz, g, ws))] if os.environ.get("WITH_FFMA") else []):
    ms, med = time_kernel(fn, 10, flush, stream)
    print(f"{shape} dil{dil} {name}: {ms*1e3:.1f} us (median {med*1e3:.1f})  {flops/ms/1e9:.1f} TFLOP/s  [incl. reduce_partials + layout permute]")
