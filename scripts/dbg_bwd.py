import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
from stereonet_b200 import ops
from stereonet_b200.autograd import functions as Fn, fused
from test_gpu_kernels import cl, uncl, rnd, DEV
from test_gpu_backward import randomize_bn
dil = 4
shape = (2, 32, 17, 29)
x = rnd(*shape, seed=1)
torch.manual_seed(5)
conv = torch.nn.Conv2d(32, 32, 3, padding=dil, dilation=dil); bn = torch.nn.BatchNorm2d(32); randomize_bn(bn, 7)
gy = rnd(*shape, seed=2)
res = {}
for backend in ("ffma", "tc3", "tc3"):
  gconv, gbn = copy.deepcopy(conv).to(DEV), copy.deepcopy(bn).to(DEV)
  fused.set_conv_backend(backend)
  xg = cl(x).requires_grad_()
  g = ops.geom(xg.shape, 3, dil=dil)
  z, stats = fused.conv3x3_c32(xg.detach(), gconv, g, bias=gconv.bias.detach(), want_stats=True)
  scale, shift, mean, invstd = ops.bn_finalize(stats, z.numel() // 32, gbn)
  dy = cl(gy)
  dz, dgamma, dbeta, dbias = ops.bn_lrelu_bwd(z, dy, scale, shift, mean, invstd, True)
  dx, _ = fused.conv3x3_c32_dgrad(dz, gconv, g, residual=dy)
  dx2, _ = fused.conv3x3_c32_dgrad(dz, gconv, g)
  torch.cuda.synchronize()
  cur = dict(z=z, scale=scale, shift=shift, mean=mean, invstd=invstd, dz=dz, dx=dx, dx2=dx2, stats=stats.double().sum(0))
  if "ffma" in res:
    for k in cur:
      a, b = res["ffma"][k], cur[k]
      print(backend, k, f"max diff {(a - b).abs().max().item():.3e} / mag {a.abs().max().item():.3e}")
  else:
    res["ffma"] = cur
