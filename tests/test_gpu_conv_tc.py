"""GPU parity of the tcgen05 (tensor-core) 32->32 convolutions against plain fp32 torch on the CPU.  Kernels / operand formats:
"ws" = snb_conv_c32_ws (the product kernel, error-compensated fp16 split), "h" = the same split on the round-1 kernels and
3 = 3xTF32 must be fp32-grade: 1e-5 of the output magnitude; 1 = plain single-pass TF32: 3e-3."""
import pytest
import torch
import torch.nn.functional as F

from stereonet_b200 import ops
from test_gpu_kernels import cl, uncl, rnd, close, DEV

pytestmark = pytest.mark.gpu
TOL = {"ws": 1e-5, "h": 1e-5, 3: 1e-5, 1: 3e-3}


def kw(fmt):
  return dict(fmt=fmt)


def wimg(w, fmt, mode=0):
  return ops.prep_conv_weights_tc(w.to(DEV), mode=mode, fmt=fmt)


@pytest.mark.parametrize("fmt", ["ws", "h", 3, 1])
@pytest.mark.parametrize("B,H,W,dil", [(1, 19, 45, 1), (2, 23, 37, 2), (1, 40, 50, 4), (1, 33, 41, 8), (1, 8, 128, 1),
                                       (1, 47, 156, 1), (1, 130, 260, 2), (2, 5, 300, 8), (1, 1, 7, 1), (1, 9, 129, 1),
                                       (1, 70, 256, 16), (3, 3, 128, 2)])
def test_conv_tc_2d(B, H, W, dil, fmt):
  """Vertical-walk 2-D kernel (snb_conv2d_c32_tc), TMA producer."""
  x, w, b = rnd(B, 32, H, W, seed=1), rnd(32, 32, 3, 3, seed=2, scale=0.1), rnd(32, seed=3)
  ref = F.conv2d(x, w, b, padding=dil, dilation=dil)
  g = ops.geom((B, H, W, 32), 3, dil=dil)
  y, _ = ops.conv_c32_tc(cl(x), wimg(w, fmt), g, bias=b.to(DEV), **kw(fmt))
  close(uncl(y), ref, TOL[fmt], f"tc conv2d dil={dil} fmt={fmt}")


@pytest.mark.parametrize("wscale,xscale", [(1e-4, 1.0), (3.0, 1.0), (0.1, 1e-3), (0.1, 300.0), (0.02, 1e-6)])
def test_conv_tc_f16_split_dynamic_range(wscale, xscale):
  """The fp16 split must stay fp32-grade whatever the magnitude of weights (power-of-two scaling in the weight image) and
  activations (the low part is pre-scaled by 2^11, so it never falls into the fp16 subnormals) — within |x| < 65504."""
  x, w, b = rnd(1, 32, 21, 150, seed=1) * xscale, rnd(32, 32, 3, 3, seed=2, scale=wscale), rnd(32, seed=3) * (wscale * xscale)
  x[0, :, 3, 5] = 0.0                                           # exact zeros and a few large outliers
  x[0, 7, 10, 20] = 6.0e4 if xscale >= 1.0 else x[0, 7, 10, 20]
  ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).float()
  for fmt in ("ws", "h"):
    y, _ = ops.conv_c32_tc(cl(x), wimg(w, fmt), ops.geom((1, 21, 150, 32), 3), bias=b.to(DEV), fmt=fmt)
    close(uncl(y), ref, 1e-5, f"f16 split [{fmt}] wscale={wscale} xscale={xscale}")
  x3, w3 = rnd(1, 32, 5, 9, 40, seed=4) * xscale, rnd(32, 32, 3, 3, 3, seed=5, scale=wscale)
  ref3 = F.conv3d(x3.double(), w3.double(), None, padding=1).float()
  y3, _ = ops.conv_c32_tc(cl(x3), wimg(w3, "ws"), ops.geom((1, 5, 9, 40, 32), 3))
  close(uncl(y3), ref3, 1e-5, f"f16 split [ws 3-D] wscale={wscale} xscale={xscale}")


@pytest.mark.parametrize("fmt", ["ws", "h", 3])
@pytest.mark.parametrize("B,H,W,dil", [(2, 23, 37, 2), (1, 47, 156, 1), (1, 64, 300, 4)])
def test_conv2d_tc_full_epilogue_and_stats(B, H, W, dil, fmt):
  x, w, b = rnd(B, 32, H, W, seed=1), rnd(32, 32, 3, 3, seed=2, scale=0.1), rnd(32, seed=3)
  scale, shift = rnd(32, seed=4).abs() + 0.5, rnd(32, seed=5)
  z = F.conv2d(x, w, b, padding=dil, dilation=dil)
  ref = F.leaky_relu(z * scale.view(1, 32, 1, 1) + shift.view(1, 32, 1, 1), 0.2) + x
  xc = cl(x)
  if fmt == "ws":     # snb_conv_c32_ws: statistics of conv + bias only (the train-mode call), the full epilogue separately
    g = ops.geom((B, H, W, 32), 3, dil=dil)
    zc, stats = ops.conv_c32_tc(xc, wimg(w, fmt), g, bias=b.to(DEV), want_stats=True)
    close(uncl(zc), z, 1e-5, "conv2d ws z")
    s = stats.double().sum(0).cpu()
    close(s[0], z.double().sum((0, 2, 3)), 1e-4, "sum z")
    close(s[1], (z.double() ** 2).sum((0, 2, 3)), 1e-4, "sum z^2")
    y, _ = ops.conv_c32_tc(xc, wimg(w, fmt), g, bias=b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), residual=xc, lrelu=True)
    close(uncl(y), ref, 1e-5, "conv2d ws epilogue, residual = input (on-chip)")
    other = rnd(B, 32, H, W, seed=9)
    y2, _ = ops.conv_c32_tc(xc, wimg(w, fmt), g, bias=b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), residual=cl(other), lrelu=True)
    close(uncl(y2), ref - x + other, 1e-5, "conv2d ws epilogue, residual from global memory")
    y3, _ = ops.conv_c32_tc(xc, wimg(w, fmt), g)
    close(uncl(y3), z - b.view(1, 32, 1, 1), 1e-5, "conv2d ws, no epilogue")
    return
  y, stats = ops.conv_c32_tc(xc, wimg(w, fmt), ops.geom((B, H, W, 32), 3, dil=dil), bias=b.to(DEV),
                             scale=scale.to(DEV), shift=shift.to(DEV), residual=xc, lrelu=True, want_stats=True, **kw(fmt))
  close(uncl(y), ref, 1e-5, "conv2d tc epilogue")
  s = stats.double().sum(0).cpu()
  close(s[0], z.double().sum((0, 2, 3)), 1e-4, "sum z")
  close(s[1], (z.double() ** 2).sum((0, 2, 3)), 1e-4, "sum z^2")


@pytest.mark.parametrize("B,D,H,W", [(1, 24, 9, 20), (2, 6, 7, 33), (1, 12, 17, 16), (1, 24, 47, 156), (1, 1, 3, 5), (2, 3, 16, 16),
                                     (1, 24, 40, 120), (1, 12, 20, 60)])
def test_conv_ws_3d(B, D, H, W):
  """snb_conv_c32_ws, 3-D: walk along the disparity axis, two accumulator rings; statistics of conv + bias, then the full
  epilogue (incl. a residual from global memory) in separate launches."""
  x, w, b = rnd(B, 32, D, H, W, seed=1), rnd(32, 32, 3, 3, 3, seed=2, scale=0.05), rnd(32, seed=3)
  scale, shift = rnd(32, seed=4).abs() + 0.5, rnd(32, seed=5)
  z = F.conv3d(x, w, b, padding=1)
  ref = F.leaky_relu(z * scale.view(1, 32, 1, 1, 1) + shift.view(1, 32, 1, 1, 1), 0.2)
  g = ops.geom((B, D, H, W, 32), 3)
  xc = cl(x)
  zc, stats = ops.conv_c32_tc(xc, wimg(w, "ws"), g, bias=b.to(DEV), want_stats=True)
  close(uncl(zc), z, 1e-5, "ws conv3d z")
  s = stats.double().sum(0).cpu()
  close(s[0], z.double().sum((0, 2, 3, 4)), 1e-4, "sum z")
  close(s[1], (z.double() ** 2).sum((0, 2, 3, 4)), 1e-4, "sum z^2")
  y, _ = ops.conv_c32_tc(xc, wimg(w, "ws"), g, bias=b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), lrelu=True)
  close(uncl(y), ref, 1e-5, "ws conv3d epilogue")
  y2, _ = ops.conv_c32_tc(xc, wimg(w, "ws"), g, bias=b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), lrelu=True, residual=xc)
  close(uncl(y2), ref + x, 1e-5, "ws conv3d epilogue + residual")


@pytest.mark.parametrize("passes", ["h", 3, 1])
@pytest.mark.parametrize("B,D,H,W", [(1, 24, 9, 20), (2, 6, 7, 33), (1, 12, 17, 16), (1, 24, 47, 156)])
def test_conv_tc_3d_full_epilogue(B, D, H, W, passes):
  x, w, b = rnd(B, 32, D, H, W, seed=1), rnd(32, 32, 3, 3, 3, seed=2, scale=0.05), rnd(32, seed=3)
  scale, shift = rnd(32, seed=4).abs() + 0.5, rnd(32, seed=5)
  z = F.conv3d(x, w, b, padding=1)
  ref = F.leaky_relu(z * scale.view(1, 32, 1, 1, 1) + shift.view(1, 32, 1, 1, 1), 0.2) + x
  g = ops.geom((B, D, H, W, 32), 3)
  xc = cl(x)
  y, stats = ops.conv_c32_tc(xc, wimg(w, passes), g, bias=b.to(DEV), scale=scale.to(DEV),
                             shift=shift.to(DEV), residual=xc, lrelu=True, want_stats=True, **kw(passes))
  close(uncl(y), ref, TOL[passes], f"tc conv3d passes={passes}")
  s = stats.double().sum(0).cpu()
  close(s[0], z.double().sum((0, 2, 3, 4)), 10 * TOL[passes], "sum z")
  close(s[1], (z.double() ** 2).sum((0, 2, 3, 4)), 10 * TOL[passes], "sum z^2")


def test_conv_tc_kitti_refinement_size():
  x, w, b = rnd(1, 32, 376, 1248, seed=1), rnd(32, 32, 3, 3, seed=2, scale=0.1), rnd(32, seed=3)
  for dil in (1, 8):
    ref = F.conv2d(x, w, b, padding=dil, dilation=dil)
    y, _ = ops.conv_c32_tc(cl(x), wimg(w, "ws"), ops.geom((1, 376, 1248, 32), 3, dil=dil), bias=b.to(DEV))
    close(uncl(y), ref, 1e-5, f"tc conv2d 376x1248 dil={dil}")


@pytest.mark.parametrize("fmt", ["ws", "h", 3])
def test_conv_tc_dgrad_weights(fmt):
  x = rnd(1, 32, 14, 27, seed=1).requires_grad_()
  w = rnd(32, 32, 3, 3, seed=2, scale=0.1)
  gy = rnd(1, 32, 14, 27, seed=3)
  F.conv2d(x, w, None, padding=2, dilation=2).backward(gy)
  dx, _ = ops.conv_c32_tc(cl(gy), wimg(w, fmt, mode=1), ops.geom((1, 14, 27, 32), 3, dil=2), **kw(fmt))
  close(uncl(dx), x.grad, 1e-5, "tc dgrad 2d")
  x = rnd(1, 32, 6, 9, 11, seed=1).requires_grad_()
  w = rnd(32, 32, 3, 3, 3, seed=2, scale=0.05)
  gy = rnd(1, 32, 6, 9, 11, seed=3)
  F.conv3d(x, w, None, padding=1).backward(gy)
  dx, _ = ops.conv_c32_tc(cl(gy), wimg(w, fmt, mode=1), ops.geom((1, 6, 9, 11, 32), 3), **kw(fmt))
  close(uncl(dx), x.grad, 1e-5, "tc dgrad 3d")


@pytest.mark.parametrize("B,H,W", [(1, 20, 44), (2, 47, 157), (1, 188, 624), (2, 9, 260), (1, 2, 2), (1, 95, 313), (1, 1, 5), (1, 3, 2)])
def test_conv5x5s2_one_launch(B, H, W):
  """downsample[1:] (nn.Conv2d(32, 32, 5, 2, 2), stereo_net.py:64-70) as ONE launch of the walk kernel over the four polyphase
  images of its input (snb_conv5x5s2_c32_ws), odd and even sizes, against fp32 torch on the CPU."""
  x, w, b = rnd(B, 32, H, W, seed=4), rnd(32, 32, 5, 5, seed=5, scale=0.05), rnd(32, seed=6)
  ref = F.conv2d(x, w, b, stride=2, padding=2)
  ph = ops.phase_split(cl(x))
  wimg5 = ops.prep_conv5x5s2_weights_ws(w.to(DEV))
  y = ops.conv5x5s2_c32_ws(ph, wimg5, bias=b.to(DEV))
  close(uncl(y), ref, 1e-5, "5x5 stride-2 one-launch")
  if H >= 2 and W >= 2:                              # the same layer from the un-split input (strided TMA views): identical arithmetic
    y2 = ops.conv5x5s2_c32_ws_x(cl(x), wimg5, bias=b.to(DEV))
    assert torch.equal(y2, y)
