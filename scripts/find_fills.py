"""Which Python call sites launch aten::fill_/zero_ kernels during one eager adaptation step? (diagnostic)"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import stereonet_b200 as S
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from bench import synthetic_pair
dev = "cuda:0"
torch.manual_seed(123)
f, s = S.FeatureExtractorNetwork(3).to(dev), S.StereoNet(3, 1, 0).to(dev)
st = AdaptStepper(f, s, make_optimizer(f, s, capturable=True), 376, 1248)
l, r = synthetic_pair(1000)
l, r = l.to(dev), r.to(dev)
for _ in range(2):
  st.step(l, r)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
  st.step(l, r)
torch.cuda.synchronize()
cnt = collections.Counter()
for ev in prof.events():
  if ev.name in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::add_", "aten::add"):
    stack = [fr for fr in (ev.stack or []) if "stereonet_b200" in fr or "torch/optim" in fr or "torch/nn/utils" in fr or "autograd" in fr]
    cnt[(ev.name, str(ev.input_shapes)[:40], " <- ".join(stack[:3])[:300])] += 1
for k, v in cnt.most_common(25):
  print(v, k)
