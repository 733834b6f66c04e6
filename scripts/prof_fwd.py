"""Two eval-mode forwards at KITTI size without the CUDA graph (for the ncu launch list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import stereonet_b200 as S
from stereonet_b200.runtime import StereoEngine
from bench import synthetic_pair
dev = "cuda:0"
torch.manual_seed(123)
f, s = S.FeatureExtractorNetwork(3).to(dev).eval(), S.StereoNet(3, 1, 0).to(dev).eval()
eng = StereoEngine(f, s, output_cost_volume=True, use_graph=False)
l, r = synthetic_pair(1000)
l, r = l.to(dev), r.to(dev)
for _ in range(int(os.environ.get("NFWD", "3"))):
  eng(l, r)
torch.cuda.synchronize()
print("ok")
