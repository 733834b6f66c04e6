timeout 300 python scripts/time_conv3d.py 2>&1 | tail -18
timeout 200 python scripts/tc_counters.py 2>&1 | grep "3d"
timeout 300 python - <<'PY'
import os, sys
sys.path.insert(0, "adaptive-stereo-icra-2021_b200"); sys.path.insert(0, ".")
import torch
from stereonet_b200 import ops
from bench import time_kernel, synthetic_pair
import stereonet_b200 as S
dev="cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev); stream = torch.cuda.current_stream()
torch.manual_seed(1)
f, s = S.FeatureExtractorNetwork(3).to(dev).eval(), S.StereoNet(3,1,0).to(dev).eval()
l, r = synthetic_pair(1000); pair = torch.cat([l, r]).to(dev)
with torch.no_grad():
  ms,_ = time_kernel(lambda: ops.conv5x5s2_c3(pair, f.downsample[0].weight, f.downsample[0].bias), 10, flush, stream); print("conv5x5s2_c3 (2 images)", round(ms*1e3,1), "us")
  coarse = torch.rand(1,47,156,device=dev)*20
  conv, bn = s.edge_aware_refinements[0].conv2d_feature[0][0], s.edge_aware_refinements[0].conv2d_feature[0][1]
  ms,_ = time_kernel(lambda: ops.refine_in_conv(coarse, l.to(dev), conv.weight, conv.bias.detach()), 10, flush, stream); print("refine_in_conv", round(ms*1e3,1), "us")
PY
