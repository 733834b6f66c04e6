"""Per-role cycle counters of the 2-D walk tensor-core conv kernel (diagnostics; snb_conv2d_c32_tc_profile)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"))
import torch
from stereonet_b200 import ops, _cabi
from stereonet_b200._cabi import ConvEpilogue
dev = "cuda:0"
torch.manual_seed(0)
names = ["prod_wait_rempty", "prod_total", "mma_wait_afull", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_total", "epi_bar",
         "epi_tmem_ld", "epi_tmem2smem", "epi_out", "cv_wait_rfull", "cv_wait_aempty", "cv_total", "cv_wait_res", "epi_bar2"]
SHAPE = tuple(int(v) for v in os.environ.get("SHAPE", "1,376,1248").split(","))
x = torch.randn(SHAPE + (32,), device=dev)
w = torch.randn(32, 32, 3, 3, device=dev) * 0.1
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
y = torch.empty_like(x)
for dil in (1, 8):
  g = ops.geom(SHAPE + (32,), 3, dil=dil)
  nt = _cabi.lib().snb_conv2d_c32_tc_num_tiles(C.byref(g))
  for fmt in ("ws",):
    wimg = ops.prep_conv_weights_tc(w, fmt=fmt)
    passes = (3 | ops.CONV_F16) if fmt == "h" else fmt
    for residual in (True, False):
      cnt = torch.zeros(148, 16, dtype=torch.int64, device=dev)
      e = ConvEpilogue(ops._p(b), ops._p(sc), ops._p(sh), ops._p(x if residual else None), None, 1)
      for _ in range(3):
        if fmt == "ws":
          _cabi.check(_cabi.lib().snb_conv_c32_ws_profile(ops._p(x), ops._p(wimg), ops._p(y), C.byref(g), C.byref(e), ops._p(cnt), ops._stream(x)), "profile")
        else:
          _cabi.check(_cabi.lib().snb_conv2d_c32_tc_profile(ops._p(x), ops._p(wimg), ops._p(y), C.byref(g), C.byref(e), passes, ops._p(cnt), ops._stream(x)), "profile")
      torch.cuda.synchronize()
      m = cnt.double().mean(0).tolist()
      print(f"dil{dil} fmt={fmt} res={int(residual)} tiles/SM={nt/148:.1f}: " + "  ".join(f"{n}={v/1e3:.2f}k" for n, v in zip(names, m)), flush=True)
