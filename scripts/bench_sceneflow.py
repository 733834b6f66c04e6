"""BASELINE.json configs[3]: SceneFlow-shaped 540x960, D = 192, batch 32 sharded over the visible ranks (one process per GPU,
no collective).  Prints one JSON line: whole-job pairs/s (device-resident inputs, CUDA-graph replay, CUDA events, max over
ranks) and the parity of sample 0 against the oracle on rank 0.  Usage: python scripts/bench_sceneflow.py   or under torchrun."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"), os.path.join(ROOT, "oracle")):
  sys.path.insert(0, p)
import torch
import torch.distributed as dist
import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200.runtime import StereoEngine
from stereonet_b200.parallel import shard_range

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
  dist.init_process_group("nccl", device_id=dev)
B, H, W, k = 32, 540, 960, 3
lo, hi = shard_range(B, rank, world)
fsd, ssd = O.make_feature_state(k, 11), O.make_stereo_state(22, sharpen=10.0)
f, s = S.FeatureExtractorNetwork(k).to(dev).eval(), S.StereoNet(k, 1, 0).to(dev).eval()
f.load_state_dict(fsd); s.load_state_dict(ssd)
pairs = [O.make_stereo_pair(1, H, W, seed=1000 + i, max_disp_px=60.0) for i in range(lo, min(hi, lo + 2))]   # 2 distinct samples, tiled
left = torch.cat([pairs[i % len(pairs)][0] for i in range(hi - lo)]).to(dev)
right = torch.cat([pairs[i % len(pairs)][1] for i in range(hi - lo)]).to(dev)
eng = StereoEngine(f, s, output_cost_volume=False)
with torch.no_grad():
  for _ in range(3):
    out = eng(left, right)
  torch.cuda.synchronize()
  if world > 1: dist.barrier()
  steps = 10
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(steps):
    out = eng(left, right)
  b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
if world > 1:
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
  ref = O.predict_disparity_left(fsd, ssd, pairs[0][0], pairs[0][1], k)
  err = (out["pred_disp_l/0"][:1].cpu() - ref["pred_disp_l/0"]).abs().max().item()
  print(json.dumps({"workload": "StereoNet k=3 D=192 fp32 forward, SceneFlow 32x3x540x960 sharded over ranks (BASELINE.json configs[3])",
                    "n_gpus": world, "pairs_per_gpu": hi - lo, "value": B * steps / (ms.item() / 1e3), "unit": "pairs/s",
                    "ms_per_batch": ms.item() / steps, "max_abs_disp_diff_vs_oracle_sample0_px": err,
                    "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9}))
if world > 1:
  dist.destroy_process_group()
