"""Out-of-bounds WRITE check without compute-sanitizer (closed on the GPU pool): the output tensor of the TMA-store / coalesced
store kernels is a view into a larger buffer pre-filled with a sentinel, and the bands before and after it must come back untouched
— at shapes whose last tile is ragged in every direction.  (scripts/gpu_sanitize.sh is the memcheck / racecheck recipe for a box
where the tool is allowed.)"""
import pytest
import torch

from stereonet_b200 import ops
from test_gpu_kernels import rnd, cl, DEV

pytestmark = pytest.mark.gpu
SENT = -12345.5
BAND = 1 << 16


def banded(shape):
  n = 1
  for d in shape:
    n *= d
  buf = torch.full((n + 2 * BAND,), SENT, device=DEV, dtype=torch.float32)
  return buf, buf[BAND:BAND + n].view(shape)


def intact(buf, n):
  return bool((buf[:BAND] == SENT).all().item()) and bool((buf[BAND + n:] == SENT).all().item())


@pytest.mark.parametrize("fmt", ["ws", "h"])
@pytest.mark.parametrize("B,H,W,dil", [(1, 7, 129, 1), (2, 5, 127, 2), (1, 33, 41, 8), (1, 1, 1, 1), (3, 3, 257, 4)])
def test_conv2d_writes_stay_inside(B, H, W, dil, fmt):
  x, w = cl(rnd(B, 32, H, W, seed=1)), rnd(32, 32, 3, 3, seed=2, scale=0.1).to(DEV)
  g = ops.geom((B, H, W, 32), 3, dil=dil)
  for residual in (None, "x", "other"):
    buf, out = banded((B, H, W, 32))
    res = None if residual is None else (x if residual == "x" else torch.ones_like(x))
    if fmt == "h" and residual is not None:
      continue
    y, _ = ops.conv_c32_tc(x, ops.prep_conv_weights_tc(w, 0, fmt=fmt), g, residual=res, fmt=fmt, out=out)
    torch.cuda.synchronize()
    assert intact(buf, out.numel()), (fmt, residual)
    assert bool((out != SENT).all().item())                 # every output position was written


@pytest.mark.parametrize("fmt", ["ws", "h"])
@pytest.mark.parametrize("B,D,H,W", [(1, 5, 7, 19), (2, 3, 9, 130), (1, 24, 6, 43), (1, 1, 1, 1)])
def test_conv3d_writes_stay_inside(B, D, H, W, fmt):
  x, w = torch.randn(B, D, H, W, 32, device=DEV), torch.randn(32, 32, 3, 3, 3, device=DEV) * 0.05
  g = ops.geom((B, D, H, W, 32), 3)
  buf, out = banded((B, D, H, W, 32))
  ops.conv_c32_tc(x, ops.prep_conv_weights_tc(w, 0, fmt=fmt), g, fmt=fmt, out=out)
  torch.cuda.synchronize()
  assert intact(buf, out.numel()), fmt
  assert bool((out != SENT).all().item())
