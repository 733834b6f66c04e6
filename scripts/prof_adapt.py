"""A few adaptation steps at KITTI size (for the ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import stereonet_b200 as S
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from bench import synthetic_pair
dev = "cuda:0"
torch.manual_seed(123)
f, s = S.FeatureExtractorNetwork(3).to(dev), S.StereoNet(3, 1, 0).to(dev)
st = AdaptStepper(f, s, make_optimizer(f, s, capturable=True), 376, 1248, 
                  use_graph=os.environ.get('GRAPH') == '1')
l, r = synthetic_pair(1000)
l, r = l.to(dev), r.to(dev)
n = int(os.environ.get("NSTEPS", "2"))
for _ in range(n):
  st.step(l, r)
torch.cuda.synchronize()
print("ok")
