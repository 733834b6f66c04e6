import os, sys
sys.path.insert(0, "adaptive-stereo-icra-2021_b200"); sys.path.insert(0, ".")
import torch
from stereonet_b200 import ops
from bench import time_kernel
dev="cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev); stream = torch.cuda.current_stream()
for B in (1, 8):
  fl = torch.randn(B, 47, 156, 32, device=dev); fr = torch.randn(B, 47, 156, 32, device=dev)
  ms,_ = time_kernel(lambda: ops.cost_volume(fl, fr, 24), 20, flush, stream)
  nb = 4*32*47*156*26*B
  # back-to-back without flush (device time per launch, launch overhead amortised)
  for _ in range(3): ops.cost_volume(fl, fr, 24)
  a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
  a.record(); 
  for _ in range(50): ops.cost_volume(fl, fr, 24)
  b.record(); torch.cuda.synchronize()
  print(f"B={B}: flushed single launch {ms*1e3:.1f} us ({nb/ms/1e6:.0f} GB/s); back-to-back {a.elapsed_time(b)/50*1e3:.1f} us ({nb/(a.elapsed_time(b)/50)/1e6:.0f} GB/s)")
