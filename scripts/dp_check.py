"""2-GPU check of the shared-model data-parallel adaptation step (graph and eager), with progress markers and a
faulthandler dump if anything stalls.  torchrun --nproc-per-node 2 scripts/dp_check.py"""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "90")), exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
  print(f"[rank {rank} t={time.time() % 1000:.1f}]", *a, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
say("pg up")
import stereonet_b200 as S
from stereonet_b200 import parallel
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from bench import synthetic_pair
H, W = 96, 256
res = {}
for use_graph in (False, True):
  torch.manual_seed(123)
  f, s = S.FeatureExtractorNetwork(3).to(dev), S.StereoNet(3, 1, 0).to(dev)
  st = AdaptStepper(f, s, make_optimizer(f, s, capturable=True), H, W, use_graph=use_graph)
  used = parallel.used_parameters(s, f)
  for i in range(3):
    l, r = synthetic_pair(1000 + 10 * i + rank, H, W)
    loss, fcs, out = st.step(l.to(dev), r.to(dev), dp_params=used)
    torch.cuda.synchronize()
    say("graph" if use_graph else "eager", "step", i, "loss", float(loss))
  flat = torch.cat([p.detach().reshape(-1) for p in used])
  other = flat.clone()
  dist.broadcast(other, src=0)
  say("replicas identical:", bool(torch.equal(flat, other)))
  res[use_graph] = flat
d = (res[False] - res[True]).abs().max().item()
say("eager vs graph max weight diff", d)
assert d < 1e-5
dist.barrier()
dist.destroy_process_group()
say("done")
