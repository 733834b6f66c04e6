"""Per-C-ABI-call GPU time of one eval forward (or one adaptation step: MODE=adapt) at KITTI size, CUDA events around every
library call (eager launches, warm caches).  A quick substitute for an ncu launch list."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")); sys.path.insert(0, ROOT)
import torch
import stereonet_b200 as S
from stereonet_b200 import _cabi
from bench import synthetic_pair
dev = "cuda:0"
torch.manual_seed(123)
lib = _cabi.lib()
records = []
stream = torch.cuda.current_stream()


def wrap(name, fn):
  def inner(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); rc = fn(*a); e1.record(stream)
    records.append((name, e0, e1))
    return rc
  return inner


class Proxy:
  def __init__(self, l):
    self._l = l
    self._w = {}
  def __getattr__(self, n):
    if n not in self._w:
      f = getattr(self._l, n)
      self._w[n] = wrap(n, f) if n in _cabi.SIGNATURES and not n.endswith(("num_tiles", "_floats", "num_partials", "num_blocks", "last_error", "version")) else f
    return self._w[n]


_cabi._lib = Proxy(lib)
f, s = S.FeatureExtractorNetwork(3).to(dev), S.StereoNet(3, 1, 0).to(dev)
l, r = synthetic_pair(1000)
l, r = l.to(dev), r.to(dev)
mode = os.environ.get("MODE", "fwd")
if mode == "fwd":
  f.eval(); s.eval()
  def run():
    with torch.no_grad():
      pair = torch.cat([l, r])
      feats = f(pair)
      return s(l, feats[:1], feats[1:], "l", output_cost_volume=True)
else:
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  f.train(); s.train()
  st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5), 376, 1248)
  def run():
    return st.step(l, r)
for _ in range(3):
  run()
torch.cuda.synchronize()
records.clear()
N = 5
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(stream)
for _ in range(N):
  run()
t1.record(stream)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, a, b in records:
  d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b) * 1e3
tot = sum(v[1] for v in agg.values()) / N
print(f"{mode}: {len(records)//N} library calls per run, sum of call times {tot:.1f} us, wall {t0.elapsed_time(t1)*1e3/N:.1f} us (eager)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print(f"{v[1]/N:9.1f} us  {v[0]//N:4d} calls  {v[1]/v[0]:7.1f} us/call  {100*v[1]/N/tot:5.1f}%  {k}")
