"""GPU parity of the tensor-core weight-gradient kernel (snb_conv_c32_wgrad_tc) against torch autograd on the CPU and
against the library's own FFMA weight-gradient kernel, through the C-ABI."""
import pytest
import torch
import torch.nn.functional as F

from stereonet_b200 import ops
from test_gpu_kernels import cl, rnd, close, DEV

pytestmark = pytest.mark.gpu


def ref_wgrad(x, dz, dil, three_d):
  w = torch.zeros((32, 32, 3, 3, 3) if three_d else (32, 32, 3, 3), requires_grad=True)
  y = F.conv3d(x, w, padding=1) if three_d else F.conv2d(x, w, padding=dil, dilation=dil)
  y.backward(dz)
  return w.grad


CASES_2D = [  # B, H, W, dil
  (1, 5, 16, 1), (2, 17, 29, 1), (1, 23, 37, 2), (1, 40, 50, 4), (1, 33, 141, 8), (1, 9, 300, 1), (2, 12, 130, 2),
  (1, 3, 8, 4), (1, 47, 156, 1),
]


@pytest.mark.parametrize("B,H,W,dil", CASES_2D)
def test_wgrad_tc_2d(B, H, W, dil):
  x, dz = rnd(B, 32, H, W, seed=1), rnd(B, 32, H, W, seed=2)
  ref = ref_wgrad(x, dz, dil, False)
  xg, dzg = cl(x), cl(dz)
  g = ops.geom(xg.shape, 3, stride=1, dil=dil)
  got = ops.conv_c32_wgrad_tc(xg, dzg, g, (32, 32, 3, 3), passes=3).cpu()
  close(got, ref, 2e-5, "dW (3xTF32)")
  got1 = ops.conv_c32_wgrad_tc(xg, dzg, g, (32, 32, 3, 3), passes=1).cpu()
  close(got1, ref, 5e-3, "dW (TF32)")


@pytest.mark.parametrize("B,D,H,W", [(1, 6, 9, 21), (2, 4, 5, 60), (1, 24, 7, 156), (1, 3, 3, 8), (1, 12, 16, 120)])
def test_wgrad_tc_3d(B, D, H, W):
  x, dz = rnd(B, 32, D, H, W, seed=1), rnd(B, 32, D, H, W, seed=2)
  ref = ref_wgrad(x, dz, 1, True)
  xg, dzg = cl(x), cl(dz)
  g = ops.geom(xg.shape, 3, stride=1, dil=1)
  got = ops.conv_c32_wgrad_tc(xg, dzg, g, (32, 32, 3, 3, 3), passes=3).cpu()
  close(got, ref, 2e-5, "dW (3xTF32)")


def test_wgrad_tc_exact_integers():
  """Small-integer data: every product and partial sum is exact in TF32/fp32, so the result must be bit-identical."""
  gen = torch.Generator().manual_seed(3)
  x = torch.randint(-3, 4, (1, 32, 21, 45), generator=gen).float()
  dz = torch.randint(-3, 4, (1, 32, 21, 45), generator=gen).float()
  for dil in (1, 2, 4):
    ref = ref_wgrad(x, dz, dil, False)
    g = ops.geom((1, 21, 45, 32), 3, stride=1, dil=dil)
    got = ops.conv_c32_wgrad_tc(cl(x), cl(dz), g, (32, 32, 3, 3), passes=1).cpu()
    assert torch.equal(got, ref), f"dil {dil}: max diff {(got - ref).abs().max().item()}"


@pytest.mark.parametrize("shape,dil", [((1, 376, 1248, 32), 1), ((1, 376, 1248, 32), 8), ((1, 24, 47, 156, 32), 1)])
def test_wgrad_tc_full_size_vs_ffma(shape, dil):
  """BASELINE.json configs[1]/[2] sizes: the two independent kernels of the library must agree (fp32 summation order
  differs: tolerance 1e-4 of the gradient magnitude)."""
  gen = torch.Generator(device=DEV).manual_seed(5)
  x = torch.randn(shape, device=DEV, generator=gen)
  dz = torch.randn(shape, device=DEV, generator=gen)
  g = ops.geom(shape, 3, stride=1, dil=dil)
  wshape = (32, 32, 3, 3, 3) if len(shape) == 5 else (32, 32, 3, 3)
  a = ops.conv_c32_wgrad_tc(x, dz, g, wshape, passes=3).cpu()
  b = ops.conv_c32_wgrad(x, dz, g, wshape).cpu()
  close(a, b, 1e-4, "tc vs ffma dW")
