"""In-graph cost of one small / medium / large ws convolution: a CUDA graph of N dependent launches (ping-pong buffers), replayed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
from stereonet_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
N = 20
for shape, dil, fmt in [((2, 47, 156, 32), 1, "ws"), ((2, 94, 312, 32), 1, "ws"), ((1, 376, 1248, 32), 1, "ws"), ((1, 376, 1248, 32), 4, "ws"),
                        ((1, 24, 47, 156, 32), 1, "h"), ((1, 24, 47, 156, 32), 1, "ws")]:
  ks = (3, 3, 3) if len(shape) == 5 else (3, 3)
  w = torch.randn(32, 32, *ks, device=dev) * (0.05 if len(shape) == 5 else 0.08)
  b = torch.randn(32, device=dev) * 0.1; sc = torch.rand(32, device=dev) * 0.2 + 0.9; sh = torch.randn(32, device=dev) * 0.1
  wimg = ops.prep_conv_weights_tc(w, fmt=fmt)
  g = ops.geom(shape, 3, dil=dil)
  bufs = [torch.randn(shape, device=dev), torch.empty(shape, device=dev)]
  def chain():
    for i in range(N):
      ops.conv_c32_tc(bufs[i & 1], wimg, g, bias=b, scale=sc, shift=sh, lrelu=True, fmt=fmt, out=bufs[(i + 1) & 1])
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    chain()
  torch.cuda.synchronize()
  gr = torch.cuda.CUDAGraph()
  with torch.cuda.graph(gr):
    chain()
  for _ in range(3):
    gr.replay()
  torch.cuda.synchronize()
  a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(10):
    gr.replay()
  e.record()
  torch.cuda.synchronize()
  print(f"{shape} dil{dil} fmt {fmt}: {a.elapsed_time(e) * 1e3 / (10 * N):7.2f} us per launch inside a graph of {N} dependent launches", flush=True)
