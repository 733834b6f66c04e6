"""The round-2 product kernels at KITTI size, three launches each (target of `ncu --set full -k regex:...`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
from stereonet_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
x2 = torch.randn(1, 376, 1248, 32, device=dev); w2 = torch.randn(32, 32, 3, 3, device=dev) * 0.1
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
x3 = torch.randn(1, 24, 47, 156, 32, device=dev); w3 = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
w1 = torch.randn(1, 32, 3, 3, 3, device=dev) * 0.5; b1 = torch.randn(1, device=dev)
fl = torch.randn(1, 47, 156, 32, device=dev); fr = torch.randn(1, 47, 156, 32, device=dev)
fl8 = torch.randn(8, 47, 156, 32, device=dev); fr8 = torch.randn(8, 47, 156, 32, device=dev)
wi2 = ops.prep_conv_weights_tc(w2, fmt="ws"); wi3h = ops.prep_conv_weights_tc(w3, fmt="h"); wi3 = ops.prep_conv_weights_tc(w3, fmt="ws")
g2 = ops.geom(tuple(x2.shape), 3, dil=1); g3 = ops.geom(tuple(x3.shape), 3)
img2 = torch.rand(2, 3, 376, 1248, device=dev); w0 = torch.randn(32, 3, 5, 5, device=dev) * 0.1
w5 = torch.randn(32, 32, 5, 5, device=dev) * 0.05; wi5 = ops.prep_conv5x5s2_weights_ws(w5)
xs = torch.randn(2, 47, 156, 32, device=dev); gs = ops.geom(tuple(xs.shape), 3)
cz = torch.rand(1, 47, 156, device=dev) * 20; wr = torch.randn(32, 4, 3, 3, device=dev) * 0.1
wt = torch.randn(1, 32, 3, 3, device=dev) * 0.1; bt = torch.randn(1, device=dev)
if os.environ.get("PROF_SET", "all") in ("b", "all"):  # the rest of the forward: first conv, P4, small layers, refinement in / out
  for _ in range(int(os.environ.get("PROF_REPS", "3"))):
    flush.zero_()
    ph = ops.conv5x5s2_c3_phases(img2, w0, b)
    flush.zero_()
    ops.conv5x5s2_c32_ws(ph, wi5, bias=b)
    flush.zero_()
    ops.conv_c32_tc(xs, wi2, gs, bias=b, scale=sc, shift=sh, residual=xs, lrelu=True, fmt="ws")
    flush.zero_()
    up, z, _ = ops.refine_in_conv(cz, img2[:1], wr, b, lrelu=True)
    flush.zero_()
    taps = ops.conv_c32_taps(x2, wt, 9)
    flush.zero_()
    ops.tapsum_refine_out(taps, bt, up)
    flush.zero_()
    ops.upsample_bilinear(cz, 376, 1248, 8.0)
  torch.cuda.synchronize()
  if os.environ.get("PROF_SET", "all") == "b":
    print("ok")
    sys.exit(0)
for _ in range(int(os.environ.get("PROF_REPS", "3"))):
  flush.zero_()
  ops.conv_c32_tc(x2, wi2, g2, bias=b, scale=sc, shift=sh, residual=x2, lrelu=True, fmt="ws")
  flush.zero_()
  ops.conv_c32_tc(x3, wi3h, g3, bias=b, scale=sc, shift=sh, lrelu=True, fmt="h")
  flush.zero_()
  ops.conv_c32_tc(x3, wi3, g3, bias=b, scale=sc, shift=sh, lrelu=True, fmt="ws")
  flush.zero_()
  ops.conv3d_out_softargmin(x3, w1, b1, True, True)
  flush.zero_()
  ops.cost_volume(fl, fr, 24)
  flush.zero_()
  ops.conv_c32_tc(x2, wi2, ops.geom(tuple(x2.shape), 3, dil=4), bias=b, scale=sc, shift=sh, residual=x2, lrelu=True, fmt="ws")
  flush.zero_()
  ops.cost_volume(fl8, fr8, 24)
  flush.zero_()
  ops.conv_c32_wgrad_tc(x2, x2, g2, (32, 32, 3, 3))
  flush.zero_()
  ops.conv_c32_wgrad_tc(x3, x3, g3, (32, 32, 3, 3, 3))
torch.cuda.synchronize()
print("ok")
