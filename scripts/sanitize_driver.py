"""Small-shape pass over every kernel family of libsnb200.so, meant to run under compute-sanitizer (scripts/gpu_sanitize.sh):
eval forward (k = 3 and k = 4), the engine's CUDA-graph replay, one adaptation step with the experience-replay term and the
fused clip + Adam update, the 3-D cross-check kernels.  Exits non-zero when a result is not finite."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stereonet_b200 as S  # noqa: E402
from stereonet_b200 import ops  # noqa: E402
from stereonet_b200.adapt import AdaptStepper, make_optimizer  # noqa: E402
from stereonet_b200.runtime import StereoEngine  # noqa: E402
from bench import synthetic_pair  # noqa: E402

dev = torch.device("cuda:0")
H, W = 64, 160
ok = True


def check(name, t):
  global ok
  fin = bool(torch.isfinite(t).all().item())
  print(f"{name}: shape {tuple(t.shape)} finite={fin}", flush=True)
  ok = ok and fin


for k in (3, 4):
  torch.manual_seed(123)
  f, s = S.FeatureExtractorNetwork(k).to(dev).eval(), S.StereoNet(k, 1, 0).to(dev).eval()
  l, r, gt = (t.to(dev) for t in synthetic_pair(5, H, W, slope=20.0))
  with torch.no_grad():
    out = s(l, f(l), f(r), "l", output_cost_volume=True)
  check(f"forward k={k}", out["pred_disp_l/0"])
  if k == 3:
    eng = StereoEngine(f, s, output_cost_volume=True, use_graph=os.environ.get("SANITIZE_GRAPH", "1") == "1")
    check("engine", eng(l, r)["pred_disp_l/0"])
    eng.synchronize()

torch.manual_seed(123)
f, s = S.FeatureExtractorNetwork(3).to(dev).train(), S.StereoNet(3, 1, 0).to(dev).train()
st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, fused=True), H, W, clip_grad_norm=True)
l2, r2, g2 = (t.to(dev) for t in synthetic_pair(6, H, W, slope=15.0))
for i in range(2):
  loss = st.step(l, r, replay=(l2, r2, g2))[0]
  check(f"adapt step {i} loss", loss.reshape(1))
check("weights after", torch.cat([p.detach().reshape(-1) for p in list(f.parameters()) + list(s.parameters())]))

# cross-check kernels kept next to the product path
x3 = torch.randn(1, 6, 9, 40, 32, device=dev)
w3 = torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05
g = ops.geom(tuple(x3.shape), 3)
for fmt in ("ws", "h", 3):
  y, _ = ops.conv_c32_tc(x3, ops.prep_conv_weights_tc(w3, 0, fmt=fmt), g, fmt=fmt)
  check(f"conv3d fmt={fmt}", y)
torch.cuda.synchronize()
print("sanitize driver:", "OK" if ok else "NON-FINITE RESULT")
sys.exit(0 if ok else 1)
