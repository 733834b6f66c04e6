"""3-D conv duration vs amount of work (fixed overhead vs per-tile cost), warm and L2-flushed."""
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
w = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
wimg = ops.prep_conv_weights_tc(w)
for shape in [(1, 3, 47, 156, 32), (1, 6, 47, 156, 32), (1, 12, 47, 156, 32), (1, 24, 47, 156, 32), (2, 24, 47, 156, 32), (4, 24, 47, 156, 32)]:
  x = torch.randn(shape, device=dev)
  g = ops.geom(shape, 3)
  fn = lambda: ops.conv_c32_tc(x, wimg, g, bias=b, scale=sc, shift=sh, lrelu=True, passes=3)
  ms, _ = time_kernel(fn, 10, flush, stream)
  for _ in range(3): fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): fn()
  e1.record(); torch.cuda.synchronize()
  tiles = shape[0] * shape[1] * ((47 * 156 + 125) // 126)
  print(f"{shape}: tiles {tiles:5d} ({tiles/148:5.1f}/SM)  flushed {ms*1e3:6.1f} us   back-to-back {e0.elapsed_time(e1)/20*1e3:6.1f} us", flush=True)
