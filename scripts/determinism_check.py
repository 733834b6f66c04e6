"""Back-to-back launches of each hot kernel on the same input must be bitwise identical (race detector)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import ops
from stereonet_b200.runtime import StereoEngine
dev = "cuda:0"
torch.manual_seed(0)
N = 24


def check(name, fn):
  outs = [fn() for _ in range(N)]
  torch.cuda.synchronize()
  bad = [i for i in range(1, N) if not all(torch.equal(a, b) for a, b in zip(outs[0], outs[i]))]
  worst = max((float((a - b).abs().max()) for i in bad for a, b in zip(outs[0], outs[i])), default=0.0)
  print(f"{name}: {'OK' if not bad else 'NONDETERMINISTIC in runs ' + str(bad[:8]) + f' max diff {worst:.3e}'}", flush=True)


x2 = torch.randn(1, 376, 1248, 32, device=dev); w2 = torch.randn(32, 32, 3, 3, device=dev) * 0.1
b = torch.randn(32, device=dev); sc = torch.rand(32, device=dev) + 0.5; sh = torch.randn(32, device=dev)
for dil in (1, 4):
  g = ops.geom(tuple(x2.shape), 3, dil=dil)
  for fmt in ("ws", "h", 3):
    wimg = ops.prep_conv_weights_tc(w2, fmt=fmt)
    check(f"conv2d dil{dil} fmt {fmt}", lambda: (ops.conv_c32_tc(x2, wimg, g, bias=b, scale=sc, shift=sh, residual=x2, lrelu=True, fmt=fmt)[0],))
x3 = torch.randn(1, 24, 47, 156, 32, device=dev); w3 = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
g3 = ops.geom(tuple(x3.shape), 3)
for fmt in ("ws", "h", 3):
  wimg3 = ops.prep_conv_weights_tc(w3, fmt=fmt)
  check(f"conv3d fmt {fmt}", lambda: (ops.conv_c32_tc(x3, wimg3, g3, bias=b, scale=sc, shift=sh, lrelu=True, fmt=fmt)[0],))
w1 = torch.randn(1, 32, 3, 3, 3, device=dev) * 0.5; b1 = torch.randn(1, device=dev)
check("head fused", lambda: tuple(t for t in ops.conv3d_out_softargmin(x3, w1, b1, True, True)))
fl = torch.randn(1, 47, 156, 32, device=dev); fr = torch.randn(1, 47, 156, 32, device=dev)
check("cost volume", lambda: (ops.cost_volume(fl, fr, 24),))

# whole forward: eager vs graph replays
f = S.FeatureExtractorNetwork(3).to(dev).eval(); s = S.StereoNet(3, 1, 0).to(dev).eval()
f.load_state_dict(O.make_feature_state(3, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=40.0))
l, r, _ = O.make_stereo_pair(1, 376, 1248, seed=1000, max_disp_px=60.0)
l, r = l.to(dev), r.to(dev)
with torch.no_grad():
  def fwd():
    o = s(l, f(l), f(r), "l", output_cost_volume=True)
    return (o["pred_disp_l/0"].clone(), o["pred_disp_l/3"].clone(), o["cost_volume_l/3"].clone())
  check("forward eager", fwd)
  eng = StereoEngine(f, s, output_cost_volume=True)
  def fwdg():
    o = eng(l, r)
    return (o["pred_disp_l/0"].clone(), o["pred_disp_l/3"].clone(), o["cost_volume_l/3"].clone())
  check("forward graph", fwdg)
  a = fwd(); bgr = fwdg()
  torch.cuda.synchronize()
  print("eager vs graph:", [float((p - q).abs().max()) for p, q in zip(a, bgr)])
