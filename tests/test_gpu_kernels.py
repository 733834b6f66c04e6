"""GPU parity tests, kernel by kernel, through the C-ABI (stereonet_b200.ops -> libsnb200.so) against the oracle /
plain fp32 torch on the CPU.  Tolerances are written per test (fp32 arithmetic, different summation order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import stereonet_oracle as O
from stereonet_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cl(x):   # NCHW / NCDHW -> channels-last contiguous on the GPU
  perm = (0, 2, 3, 1) if x.dim() == 4 else (0, 2, 3, 4, 1)
  return x.permute(*perm).contiguous().to(DEV)


def uncl(y):
  perm = (0, 3, 1, 2) if y.dim() == 4 else (0, 4, 1, 2, 3)
  return y.permute(*perm).cpu()


def rnd(*shape, seed=0, scale=1.0):
  return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(got, ref, rtol, what=""):
  err = (got - ref).abs().max().item()
  mag = ref.abs().max().item() + 1e-12
  assert err <= rtol * mag, f"{what}: max err {err:.3e} vs magnitude {mag:.3e} (rtol {rtol})"


@pytest.mark.parametrize("B,H,W,D", [(1, 5, 40, 24), (2, 7, 16, 24), (1, 47, 156, 24), (1, 3, 9, 12), (1, 4, 33, 6),
                                     (4, 47, 156, 24), (3, 40, 150, 40)])      # the last two (> 64 MB) take the TMA-staged kernel
def test_cost_volume_bit_exact(B, H, W, D):
  L, R = rnd(B, 32, H, W, seed=1), rnd(B, 32, H, W, seed=2)
  ref = O.cost_volume_numpy(L.numpy(), R.numpy(), D)
  got = uncl(ops.cost_volume(cl(L), cl(R), D)).numpy()
  assert np.array_equal(got, ref)        # a single fp32 subtraction per element: must be bit-exact


def test_cost_volume_bwd_matches_autograd():
  B, H, W, D = 2, 5, 37, 24
  L, R = rnd(B, 32, H, W, seed=1).requires_grad_(), rnd(B, 32, H, W, seed=2).requires_grad_()
  g = rnd(B, 32, D, H, W, seed=3)
  O.cost_volume(L, R, D).backward(g)
  dl, dr = ops.cost_volume_bwd(cl(g))
  close(uncl(dl), L.grad, 1e-6, "dleft")
  close(uncl(dr), R.grad, 1e-6, "dright")


CONV2D_CASES = [  # (B, H, W, ksize, stride, dil)
  (1, 19, 45, 3, 1, 1), (2, 23, 37, 3, 1, 2), (1, 40, 50, 3, 1, 4), (1, 33, 41, 3, 1, 8),
  (1, 47, 156, 5, 2, 1), (1, 135, 31, 5, 2, 1), (1, 8, 128, 3, 1, 1),
]


@pytest.mark.parametrize("B,H,W,ks,stride,dil", CONV2D_CASES)
def test_conv_c32_2d_plain(B, H, W, ks, stride, dil):
  x, w, b = rnd(B, 32, H, W, seed=1), rnd(32, 32, ks, ks, seed=2, scale=0.1), rnd(32, seed=3)
  pad = dil * (ks - 1) // 2
  ref = F.conv2d(x, w, b, stride=stride, padding=pad, dilation=dil)
  g = ops.geom((B, H, W, 32), ks, stride=stride, dil=dil, pad=pad)
  y, _ = ops.conv_c32(cl(x), ops.prep_conv_weights(w.to(DEV)), g, bias=b.to(DEV))
  assert tuple(uncl(y).shape) == tuple(ref.shape)
  close(uncl(y), ref, 1e-5, "conv2d")


@pytest.mark.parametrize("B,D,H,W", [(1, 24, 9, 20), (2, 6, 7, 33), (1, 12, 17, 16)])
def test_conv_c32_3d_full_epilogue(B, D, H, W):
  x, w, b = rnd(B, 32, D, H, W, seed=1), rnd(32, 32, 3, 3, 3, seed=2, scale=0.05), rnd(32, seed=3)
  scale, shift = rnd(32, seed=4).abs() + 0.5, rnd(32, seed=5)
  z = F.conv3d(x, w, b, padding=1)
  ref = F.leaky_relu(z * scale.view(1, 32, 1, 1, 1) + shift.view(1, 32, 1, 1, 1), 0.2) + x
  g = ops.geom((B, D, H, W, 32), 3)
  xc = cl(x)
  y, stats = ops.conv_c32(xc, ops.prep_conv_weights(w.to(DEV)), g, bias=b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV),
                          residual=xc, lrelu=True, want_stats=True)
  close(uncl(y), ref, 1e-5, "conv3d+epilogue")
  s = stats.double().sum(0).cpu()
  close(s[0], z.double().sum((0, 2, 3, 4)), 1e-5, "sum z")
  close(s[1], (z.double() ** 2).sum((0, 2, 3, 4)), 1e-5, "sum z^2")


def test_conv_c32_dgrad_weights():
  """The same kernel with mode-1 (flipped, transposed) weights is the data gradient of a stride-1 conv."""
  for dil in (1, 2):
    x = rnd(1, 32, 14, 27, seed=1).requires_grad_()
    w = rnd(32, 32, 3, 3, seed=2, scale=0.1)
    gy = rnd(1, 32, 14, 27, seed=3)
    F.conv2d(x, w, None, padding=dil, dilation=dil).backward(gy)
    g = ops.geom((1, 14, 27, 32), 3, dil=dil)
    dx, _ = ops.conv_c32(cl(gy), ops.prep_conv_weights(w.to(DEV), mode=1), g)
    close(uncl(dx), x.grad, 1e-5, f"dgrad dil={dil}")
  x = rnd(1, 32, 6, 9, 11, seed=1).requires_grad_()
  w = rnd(32, 32, 3, 3, 3, seed=2, scale=0.05)
  gy = rnd(1, 32, 6, 9, 11, seed=3)
  F.conv3d(x, w, None, padding=1).backward(gy)
  dx, _ = ops.conv_c32(cl(gy), ops.prep_conv_weights(w.to(DEV), mode=1), ops.geom((1, 6, 9, 11, 32), 3))
  close(uncl(dx), x.grad, 1e-5, "dgrad 3d")


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 37, 75), (1, 376, 1248)])
def test_conv5x5s2_c3(B, H, W):
  img, w, b = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(1)), rnd(32, 3, 5, 5, seed=2, scale=0.2), rnd(32, seed=3)
  ref = F.conv2d(img, w, b, stride=2, padding=2)
  y = ops.conv5x5s2_c3(img.to(DEV), w.to(DEV), b.to(DEV))
  close(uncl(y), ref, 1e-5, "conv5x5s2_c3")


@pytest.mark.parametrize("B,h,w,H,W", [(1, 8, 16, 64, 128), (2, 9, 15, 68, 120), (1, 47, 156, 376, 1248)])
def test_refine_in_conv(B, h, w, H, W):
  coarse = torch.rand(B, h, w, generator=torch.Generator().manual_seed(1)) * 20
  rgb = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(2))
  wt, b = rnd(32, 4, 3, 3, seed=3, scale=0.2), rnd(32, seed=4)
  up_ref = F.interpolate(coarse.unsqueeze(1), size=(H, W), mode="bilinear", align_corners=False) * (W / w)
  z_ref = F.conv2d(torch.cat([up_ref, rgb], 1), wt, b, padding=1)
  up, z, stats = ops.refine_in_conv(coarse.to(DEV), rgb.to(DEV), wt.to(DEV), b.to(DEV), want_stats=True)
  close(up.cpu(), up_ref.squeeze(1), 2e-6, "upsampled disparity")
  close(uncl(z), z_ref, 3e-6, "refine_in conv")
  s = stats.double().sum(0).cpu()
  close(s[0], z_ref.double().sum((0, 2, 3)), 1e-5, "sum")
  close(s[1], (z_ref.double() ** 2).sum((0, 2, 3)), 1e-5, "sumsq")
  # fused eval epilogue
  scale, shift = rnd(32, seed=5).abs() + 0.5, rnd(32, seed=6)
  _, y, _ = ops.refine_in_conv(coarse.to(DEV), rgb.to(DEV), wt.to(DEV), b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), lrelu=True)
  close(uncl(y), F.leaky_relu(z_ref * scale.view(1, 32, 1, 1) + shift.view(1, 32, 1, 1), 0.2), 3e-6, "refine_in eval")


@pytest.mark.parametrize("B,h,w,H,W", [(1, 5, 9, 40, 72), (2, 4, 17, 27, 130), (1, 12, 40, 96, 320), (1, 1, 1, 3, 5), (1, 3, 33, 20, 259)])
def test_refine_in_conv_ws(B, h, w, H, W):
  """The same layer on the walk kernel (snb_refine_pack_input + snb_conv_c4_ws: four-channel rows, one K = 16 slice per row)."""
  coarse = torch.rand(B, h, w, generator=torch.Generator().manual_seed(1)) * 20
  rgb = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(2))
  wt, b = rnd(32, 4, 3, 3, seed=3, scale=0.2), rnd(32, seed=4)
  up_ref = F.interpolate(coarse.unsqueeze(1), size=(H, W), mode="bilinear", align_corners=False) * (W / w)
  z_ref = F.conv2d(torch.cat([up_ref, rgb], 1), wt, b, padding=1)
  wimg = ops.refine_in_weights_ws(wt.to(DEV))
  up, z, stats = ops.refine_in_conv_ws(coarse.to(DEV), rgb.to(DEV), wimg, b.to(DEV), want_stats=True)
  close(up.cpu(), up_ref.squeeze(1), 2e-6, "upsampled disparity")
  close(uncl(z), z_ref, 1e-5, "refine_in conv (ws)")
  s = stats.double().sum(0).cpu()
  close(s[0], z_ref.double().sum((0, 2, 3)), 1e-5, "sum")
  close(s[1], (z_ref.double() ** 2).sum((0, 2, 3)), 1e-5, "sumsq")
  scale, shift = rnd(32, seed=5).abs() + 0.5, rnd(32, seed=6)
  _, y, _ = ops.refine_in_conv_ws(coarse.to(DEV), rgb.to(DEV), wimg, b.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV), lrelu=True)
  close(uncl(y), F.leaky_relu(z_ref * scale.view(1, 32, 1, 1) + shift.view(1, 32, 1, 1), 0.2), 1e-5, "refine_in eval (ws)")


@pytest.mark.parametrize("B,D,H,W,sharp", [(1, 24, 9, 20, 1.0), (2, 12, 7, 33, 30.0), (1, 6, 5, 70, 30.0), (1, 24, 47, 156, 10.0)])
def test_conv3d_out_softargmin(B, D, H, W, sharp):
  x, w, b = rnd(B, 32, D, H, W, seed=1), rnd(1, 32, 3, 3, 3, seed=2, scale=0.05 * sharp), rnd(1, seed=3)
  cost_ref = F.conv3d(x, w, b, padding=1).squeeze(1)
  pred_ref = O.soft_argmin(cost_ref)
  taps = ops.conv_c32_taps(cl(x), w.to(DEV), 27)
  cost, pred = ops.tapsum_softargmin(taps, b.to(DEV), True)
  close(cost.cpu(), cost_ref, 1e-5, "cost")
  assert (pred.cpu() - pred_ref).abs().max().item() <= 2e-4, "soft-argmin (coarse px)"
  _, pred2 = ops.tapsum_softargmin(taps, b.to(DEV), False)
  assert torch.equal(pred, pred2)


@pytest.mark.parametrize("B,D,H,W,sharp", [(1, 24, 9, 20, 1.0), (2, 12, 7, 33, 30.0), (1, 6, 5, 70, 30.0), (1, 24, 47, 156, 10.0),
                                           (1, 24, 40, 120, 10.0), (1, 12, 20, 60, 20.0), (1, 24, 1, 9, 10.0), (1, 24, 2, 3, 10.0),
                                           (1, 48, 11, 50, 5.0), (3, 24, 68, 120, 10.0), (1, 24, 5, 8, 30.0)])
def test_conv3d_out_softargmin_fused(B, D, H, W, sharp):
  """The ONE-kernel head (snb_conv3d_out_softargmin: conv3d_alone + softmax + expectation + cost + FCS, stereo_net.py:187-198)
  against fp32 torch / the oracle's soft-argmin / the sort-based feature-contrast score (feature_contrast.py:12-23)."""
  x, w, b = rnd(B, 32, D, H, W, seed=1), rnd(1, 32, 3, 3, 3, seed=2, scale=0.05 * sharp), rnd(1, seed=3)
  cost_ref = F.conv3d(x, w, b, padding=1).squeeze(1)
  pred_ref = O.soft_argmin(cost_ref)
  s = torch.sort(cost_ref, dim=1, descending=True)[0]
  fcs_ref = s[:, 0] - s[:, 2:].mean(dim=1)
  cost, pred, fcs = ops.conv3d_out_softargmin(cl(x), w.to(DEV), b.to(DEV), want_cost=True, want_fcs=True)
  close(cost.cpu(), cost_ref, 1e-5, "cost")
  assert (pred.cpu() - pred_ref).abs().max().item() <= 2e-4, "soft-argmin (coarse px)"
  close(fcs.cpu(), fcs_ref, 2e-5, "feature contrast")
  _, pred2, none = ops.conv3d_out_softargmin(cl(x), w.to(DEV), b.to(DEV), want_cost=False)
  assert none is None and torch.equal(pred, pred2)


@pytest.mark.parametrize("B,H,W", [(1, 20, 33), (2, 64, 300)])
def test_refine_out(B, H, W):
  x, w, b = rnd(B, 32, H, W, seed=1), rnd(1, 32, 3, 3, seed=2, scale=0.1), rnd(1, seed=3)
  up = rnd(B, H, W, seed=4)
  ref = F.relu(up.unsqueeze(1) + F.conv2d(x, w, b, padding=1)).squeeze(1)
  taps = ops.conv_c32_taps(cl(x), w.to(DEV), 9)
  out = ops.tapsum_refine_out(taps, b.to(DEV), up.to(DEV))
  close(out.cpu(), ref, 3e-6, "refine out")


@pytest.mark.parametrize("h,w,H,W", [(8, 16, 64, 128), (9, 15, 68, 120), (47, 156, 376, 1248), (5, 7, 5, 7)])
def test_upsample_fwd_bwd(h, w, H, W):
  x = rnd(2, h, w, seed=1).requires_grad_()
  ref = 8.0 * F.interpolate(x.unsqueeze(1), size=(H, W), mode="bilinear", align_corners=False).squeeze(1)
  got = ops.upsample_bilinear(x.detach().to(DEV), H, W, 8.0)
  close(got.cpu(), ref.detach(), 2e-6, "upsample")
  g = rnd(2, H, W, seed=2)
  ref.backward(g)
  din = ops.upsample_bilinear_bwd(g.to(DEV), h, w, 8.0)
  close(din.cpu(), x.grad, 1e-5, "upsample adjoint")


def test_bn_train_pieces():
  B, H, W = 2, 23, 37
  x, w, b = rnd(B, 32, H, W, seed=1), rnd(32, 32, 3, 3, seed=2, scale=0.1), rnd(32, seed=3)
  bn = torch.nn.BatchNorm2d(32)
  with torch.no_grad():
    bn.weight.copy_(rnd(32, seed=4).abs() + 0.5); bn.bias.copy_(rnd(32, seed=5))
    bn.running_mean.copy_(rnd(32, seed=6)); bn.running_var.copy_(rnd(32, seed=7).abs() + 0.5)
  import copy
  bn_gpu = copy.deepcopy(bn).to(DEV)
  bn.train()
  z = F.conv2d(x, w, b, padding=1)
  ref = x + F.leaky_relu(bn(z), 0.2)
  xc = cl(x)
  zc, stats = ops.conv_c32(xc, ops.prep_conv_weights(w.to(DEV)), ops.geom((B, H, W, 32), 3), bias=b.to(DEV), want_stats=True)
  scale, shift, mean, invstd = ops.bn_finalize(stats, B * H * W, bn_gpu)
  y = ops.bn_apply(zc, scale, shift, residual=xc, lrelu=True)
  close(uncl(y), ref.detach(), 5e-6, "bn train fwd")
  close(bn_gpu.running_mean.cpu(), bn.running_mean, 1e-6, "running_mean")
  close(bn_gpu.running_var.cpu(), bn.running_var, 1e-6, "running_var")
  close(mean.cpu(), z.mean((0, 2, 3)), 1e-5, "batch mean")
  assert int(bn_gpu.num_batches_tracked) == 1


def test_feature_contrast_matches_sort():
  """snb_feature_contrast vs the reference formulation sorted_desc[0] - mean(sorted_desc[2:])."""
  cost = rnd(2, 24, 13, 37, seed=3) * 5
  cost[0, 3, 2, 5] = cost[0, 7, 2, 5] = cost[0].max() + 1.0        # a tie for the maximum
  s = torch.sort(cost, dim=1, descending=True)[0]
  ref = s[:, 0] - s[:, 2:].mean(dim=1)
  got = ops.feature_contrast(cost.to(DEV)).cpu()
  close(got, ref, 1e-5, "feature contrast")
