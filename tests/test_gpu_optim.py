"""optim.FusedAdamClip (snb_multi_gather + snb_adam_clip_step: clip_grad_norm_ on the stereo_net group + Adam in three launches over
one flat bucket) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam — the reference's adapt.py:208-210,391-393 — on identical
gradients, incl. the checkpoint format; and the 2-GPU shared-model data-parallel step (skipped below two devices)."""
import os
import subprocess
import sys

import pytest
import torch

import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from test_gpu_kernels import DEV

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nets():
  f = S.FeatureExtractorNetwork(3).to(DEV); s = S.StereoNet(3, 1, 0).to(DEV)
  f.load_state_dict(O.make_feature_state(3, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=10.0))
  return f, s


@pytest.mark.parametrize("gscale", [1.0, 40.0])       # 40: the stereo_net gradient norm exceeds 1 -> the clip is active
def test_fused_adam_clip_matches_torch(gscale):
  fa, sa = _nets(); fb, sb = _nets()
  ref = torch.optim.Adam([{"params": sa.parameters()}, {"params": fa.parameters()}], lr=5e-5)
  opt = make_optimizer(fb, sb, lr=5e-5, fused=True)
  g = torch.Generator(device=DEV).manual_seed(1)
  for step in range(4):
    for (na, pa), (nb, pb) in zip(list(sa.named_parameters()) + list(fa.named_parameters()), list(sb.named_parameters()) + list(fb.named_parameters())):
      if ".conv2." in na:
        continue                                      # never receives a gradient (stereo_net.py:40,44-51)
      gr = torch.randn(pa.shape, device=DEV, generator=g) * gscale * (0.01 if step % 2 else 0.002)
      pa.grad = gr.clone(); pb.grad = gr.clone()
    norm_ref = torch.nn.utils.clip_grad_norm_(sa.parameters(), 1.0)
    ref.step()
    opt.step()
    assert abs(opt.grad_norm().item() - norm_ref.item()) <= 1e-5 * norm_ref.item()
    if gscale > 1:
      assert norm_ref.item() > 1.0
  for (n, pa), (_, pb) in zip(list(sa.named_parameters()) + list(fa.named_parameters()), list(sb.named_parameters()) + list(fb.named_parameters())):
    # the two implementations differ by rounding in m / (sqrt(v) + eps): a couple of ulps of the weight (measured: 1 ulp)
    assert (pa - pb).abs().max().item() <= 3e-7 * pa.abs().max().item() + 1e-9, n
  # checkpoint format: torch.optim.Adam can load ours and vice versa (adam.pth, train.py:136-137)
  sd = opt.state_dict()
  ref2 = torch.optim.Adam([{"params": sb.parameters()}, {"params": fb.parameters()}], lr=1e-3)
  ref2.load_state_dict(sd)
  assert ref2.param_groups[0]["lr"] == 5e-5
  rsd = ref.state_dict()
  assert sorted(sd["state"].keys()) == sorted(rsd["state"].keys())
  for k in rsd["state"]:
    assert float(sd["state"][k]["step"]) == float(rsd["state"][k]["step"]) == 4.0
    assert (sd["state"][k]["exp_avg"] - rsd["state"][k]["exp_avg"]).abs().max().item() <= 1e-6 * rsd["state"][k]["exp_avg"].abs().max().item() + 1e-12
    assert (sd["state"][k]["exp_avg_sq"] - rsd["state"][k]["exp_avg_sq"]).abs().max().item() <= 1e-6 * rsd["state"][k]["exp_avg_sq"].abs().max().item() + 1e-20
  opt2 = make_optimizer(fb, sb, lr=1e-3, fused=True)
  opt2.load_state_dict(rsd)
  assert (opt2.exp_avg - opt.exp_avg).abs().max().item() <= 1e-6 * opt.exp_avg.abs().max().item() and float(opt2.step_t) == 4.0 and opt2.param_groups[0]["lr"] == 5e-5


@pytest.mark.parametrize("use_graph", [False, True])
def test_adapt_step_with_fused_optimizer_matches_torch_optimizer(use_graph):
  """Three adaptation steps (adapt.py:313-337,381-394) with FusedAdamClip leave the weights torch's clip + Adam leave."""
  H, W = 96, 256
  frames = [O.make_stereo_pair(1, H, W, seed=1000 + i, max_disp_px=40.0)[:2] for i in range(3)]
  res = []
  for fused in (False, True):
    f, s = _nets()
    st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, capturable=True, fused=fused), H, W, use_graph=use_graph)
    losses = [st.step(l.to(DEV), r.to(DEV))[0].item() for l, r in frames]
    torch.cuda.synchronize()
    res.append((losses, {n: v.detach().cpu().clone() for n, v in list(s.state_dict().items()) + list(f.state_dict().items())}))
  (la, wa), (lb, wb) = res
  assert abs(la[0] - lb[0]) < 1e-6 and max(abs(a - b) for a, b in zip(la, lb)) < 3e-4, (la, lb)
  for n in wa:
    if "num_batches_tracked" in n:
      assert torch.equal(wa[n], wb[n]), n
    else:
      assert (wa[n] - wb[n]).abs().max().item() <= 3 * 2.1 * 5e-5 + 1e-5 * wa[n].abs().max().item(), n


DP_SNIPPET = r'''
import os, sys
import torch, torch.distributed as dist
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"), os.path.join(ROOT, "oracle")):
  sys.path.insert(0, p)
import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import parallel
from stereonet_b200.adapt import AdaptStepper, make_optimizer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
H, W = 96, 256
for use_graph in (False, True):
  f = S.FeatureExtractorNetwork(3).to(dev); s = S.StereoNet(3, 1, 0).to(dev)
  f.load_state_dict(O.make_feature_state(3, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=10.0))
  opt = make_optimizer(f, s, lr=5e-5, fused=True)
  bucket = parallel.DPBucket(opt, s, f)
  st = AdaptStepper(f, s, opt, H, W, use_graph=use_graph)
  for i in range(3):
    l, r, gt = O.make_stereo_pair(1, H, W, seed=1000 + 10 * i + rank, max_disp_px=40.0)      # a different stream per rank
    rl, rr, rgt = O.make_stereo_pair(1, H, W, seed=5000 + 10 * i + rank, max_disp_px=30.0)   # and a different replay sample
    st.step(l.to(dev), r.to(dev), replay=(rl.to(dev), rr.to(dev), rgt.to(dev)), dp_bucket=bucket)
  torch.cuda.synchronize()
  # replicas must be IDENTICAL, parameters and buffers (BatchNorm running statistics travel in the same all-reduce)
  state = torch.cat([t.detach().float().reshape(-1) for net in (s, f) for t in list(net.parameters()) + list(net.buffers())])
  gathered = [torch.empty_like(state) for _ in range(world)]
  dist.all_gather(gathered, state)
  for g in gathered[1:]:
    assert torch.equal(g, gathered[0]), ("replicas diverged", use_graph, float((g - gathered[0]).abs().max()))
  init = torch.cat([t.float().reshape(-1) for t in list(O.make_stereo_state(22, sharpen=10.0).values()) + list(O.make_feature_state(3, 11).values())])
  assert s.filter[0][0][1].num_batches_tracked.item() == 6          # two train-mode passes per step (frame + replay sample)
  assert not torch.equal(s.filter[0][0][0].weight.detach().cpu(), O.make_stereo_state(22, sharpen=10.0)["filter.0.0.0.weight"])
dist.barrier()
dist.destroy_process_group()
if rank == 0:
  print("DP_OK")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_shared_model_dp_keeps_replicas_identical_including_buffers():
  """SURVEY.md section 8e / BASELINE.json configs[4]: two ranks, one stream each with experience replay, ONE NCCL all-reduce per
  step over [gradients | BN running statistics]; eager and CUDA-graph steps."""
  env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571", WORLD_SIZE="2")
  procs = [subprocess.Popen([sys.executable, "-c", DP_SNIPPET, ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True) for r in range(2)]
  outs = []
  for p in procs:
    try:
      out, _ = p.communicate(timeout=600)
    except subprocess.TimeoutExpired:
      for q in procs:
        q.kill()
      raise
    outs.append(out)
  assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
  assert "DP_OK" in outs[0]
