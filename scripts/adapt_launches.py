"""Three eager adaptation steps at KITTI size with the fused optimizer (target of the ncu launch-list pass:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/adapt_launches.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import stereonet_b200 as S
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from bench import synthetic_pair
dev = torch.device("cuda:0")
torch.manual_seed(123)
f, s = S.FeatureExtractorNetwork(3).to(dev).train(), S.StereoNet(3, 1, 0).to(dev).train()
st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, fused=True), 376, 1248, clip_grad_norm=True, use_graph=False)
l, r, _ = (t.to(dev) for t in synthetic_pair(1000))
for i in range(int(os.environ.get("STEPS", "3"))):
  st.step(l, r)
torch.cuda.synchronize()
print("ok")
