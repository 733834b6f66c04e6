"""Delta-input probes of the tensor-core wgrad kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
from stereonet_b200 import ops
dev = "cuda:0"
def probe(H, W, dil, xs, zs, passes=1):
  x = torch.zeros(1, H, W, 32, device=dev); dz = torch.zeros(1, H, W, 32, device=dev)
  for (y, q, c, v) in xs: x[0, y, q, c] = v
  for (y, q, c, v) in zs: dz[0, y, q, c] = v
  g = ops.geom(x.shape, 3, stride=1, dil=dil)
  import ctypes as C
  from stereonet_b200 import _cabi
  n = _cabi.lib().snb_conv_c32_wgrad_tc_num_partials(C.byref(g))
  part = torch.zeros((n, 9 * 1024), device=dev); dbg = torch.full((n, 64), -7.0, device=dev)
  rc = _cabi.lib().snb_conv_c32_wgrad_tc_debug(ops._p(x), ops._p(dz), ops._p(part), C.byref(g), passes, ops._p(dbg), ops._stream(x))
  torch.cuda.synchronize()
  print("rc", rc, "n", n, "dbg[0][:16]", dbg[0, :16].tolist())
  print("part nnz", part.count_nonzero().item())
  got = ops.conv_c32_wgrad_tc(x, dz, g, (32, 32, 3, 3), passes=passes).cpu()
  ref = ops.conv_c32_wgrad(x, dz, g, (32, 32, 3, 3)).cpu()
  torch.cuda.synchronize()
  nz = got.nonzero().tolist()
  print(f"H{H} W{W} dil{dil} x{xs} dz{zs}: got nnz {len(nz)} nan {torch.isnan(got).sum().item()} max {got.abs().max().item():.3g}")
  for i in nz[:12]: print("    got[co,ci,kh,kw] =", i, got[tuple(i)].item())
  for i in ref.nonzero().tolist()[:6]: print("    ref[co,ci,kh,kw] =", i, ref[tuple(i)].item())
probe(5, 16, 1, [(2, 5, 3, 1.0)], [(2, 5, 7, 1.0)])
probe(5, 16, 1, [(2, 6, 3, 1.0)], [(2, 5, 7, 1.0)])
probe(5, 16, 1, [(3, 5, 3, 1.0)], [(2, 5, 7, 1.0)])
probe(5, 16, 1, [(2, 0, 0, 1.0)], [(2, 0, 0, 1.0)])
probe(5, 16, 1, [(0, 0, 0, 1.0)], [(0, 0, 0, 1.0)])
probe(5, 16, 1, [(2, 5, 3, 2.0), (2, 6, 4, 3.0)], [(2, 5, 7, 1.0), (2, 6, 8, 5.0)])
probe(5, 16, 1, [(2, 5, 3, 1.0)], [(2, 5, 7, 1.0)], passes=3)
probe(5, 16, 2, [(2, 5, 3, 1.0)], [(2, 5, 7, 1.0)])
