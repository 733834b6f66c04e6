"""Time the two small-Cin first layers (downsample[0] and the refinement input conv) at KITTI size."""
import os, sys
sys.path.insert(0, "adaptive-stereo-icra-2021_b200"); sys.path.insert(0, ".")
import torch
from stereonet_b200 import ops
from bench import time_kernel
dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev); stream = torch.cuda.current_stream()
g = torch.Generator(device=dev).manual_seed(1)
img = torch.rand(2, 3, 376, 1248, device=dev, generator=g)
w5 = torch.randn(32, 3, 5, 5, device=dev, generator=g) * 0.2; b5 = torch.randn(32, device=dev, generator=g)
ms, _ = time_kernel(lambda: ops.conv5x5s2_c3(img, w5, b5), 20, flush, stream)
print(f"conv5x5s2_c3 (2 images): {ms*1e3:.1f} us")
ms, _ = time_kernel(lambda: ops.conv5x5s2_c3_phases(img, w5, b5), 20, flush, stream)
print(f"conv5x5s2_c3_phases (2 images): {ms*1e3:.1f} us")
coarse = torch.rand(1, 47, 156, device=dev, generator=g) * 20
rgb = img[:1].contiguous()
w3 = torch.randn(32, 4, 3, 3, device=dev, generator=g) * 0.2; b3 = torch.randn(32, device=dev, generator=g)
sc = torch.rand(32, device=dev, generator=g) + 0.5; sh = torch.randn(32, device=dev, generator=g)
ms, _ = time_kernel(lambda: ops.refine_in_conv(coarse, rgb, w3, b3, scale=sc, shift=sh, lrelu=True), 20, flush, stream)
print(f"refine_in_conv eval: {ms*1e3:.1f} us ({60.1/ms/1e3*1e3:.0f} GB/s of the 60 MB output)")
ms, _ = time_kernel(lambda: ops.refine_in_conv(coarse, rgb, w3, b3, want_stats=True), 20, flush, stream)
print(f"refine_in_conv train(stats): {ms*1e3:.1f} us")
