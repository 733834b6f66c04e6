"""Time the 32->1 tap contraction (conv3d_alone / conv2d_out) at KITTI size."""
import os, sys
sys.path.insert(0, "adaptive-stereo-icra-2021_b200"); sys.path.insert(0, ".")
import torch
from stereonet_b200 import ops
from bench import time_kernel
dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev); stream = torch.cuda.current_stream()
g = torch.Generator(device=dev).manual_seed(1)
x3 = torch.randn(1, 24, 47, 156, 32, device=dev, generator=g); w3 = torch.randn(1, 32, 3, 3, 3, device=dev, generator=g) * 0.1
ms, _ = time_kernel(lambda: ops.conv_c32_taps(x3, w3, 27), 20, flush, stream)
print(f"taps27 (head): {ms*1e3:.1f} us  ({(x3.numel()*4 + 27*24*47*156*4)/ms/1e6:.0f} GB/s read+write)")
x2 = torch.randn(1, 376, 1248, 32, device=dev, generator=g); w2 = torch.randn(1, 32, 3, 3, device=dev, generator=g) * 0.1
ms, _ = time_kernel(lambda: ops.conv_c32_taps(x2, w2, 9), 20, flush, stream)
print(f"taps9 (refine out): {ms*1e3:.1f} us  ({(x2.numel()*4 + 9*376*1248*4)/ms/1e6:.0f} GB/s read+write)")
