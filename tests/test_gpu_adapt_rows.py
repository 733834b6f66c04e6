"""GPU parity of SURVEY.md section 8 rows f3 / f4: the Khamis replay loss, the two-pass experience-replay step, the device-resident
evaluation metrics and the batched OVS validation.  Pinned twice: against tests/golden/rows_f3_f4.npz, which oracle/gen_golden.py
produced by executing the reference's OWN functions (loss_functions.khamis_robust_loss, the evaluate() body of train.py:98-110,
the StateMachine.validate loop of adapt.py:122-142, one ER Adam step adapt.py:328-349,381-394), and against the oracle's
restatement of those formulas (oracle/stereonet_oracle.py) on further shapes.  The product package holds no PyTorch formulation."""
import os

import numpy as np
import pytest
import torch

import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import ops
from stereonet_b200.losses import khamis_robust_loss, monodepth_single_loss, feature_contrast_mean
from test_gpu_kernels import DEV

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rows_f3_f4.npz"))


def _nets(k=3, sharpen=10.0):
  f = S.FeatureExtractorNetwork(k).to(DEV); s = S.StereoNet(k, 1, 0).to(DEV)
  f.load_state_dict(O.make_feature_state(k, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=sharpen))
  return f, s


def test_khamis_loss_vs_reference_golden():
  """snb_khamis_loss against the value and gradient the reference's khamis_robust_loss produced (gen_golden.py)."""
  pred = torch.from_numpy(GOLD["khamis/pred"]).to(DEV).requires_grad_()
  gt = torch.from_numpy(GOLD["khamis/gt"]).to(DEV)
  loss = khamis_robust_loss(pred, gt)
  (3.0 * loss).backward()
  assert abs(loss.item() - float(GOLD["khamis/loss"])) <= 2e-6 * max(1.0, float(GOLD["khamis/loss"]))
  gref = torch.from_numpy(GOLD["khamis/dpred"])
  assert (pred.grad.cpu() / 3.0 - gref).abs().max().item() <= 1e-6 * gref.abs().max().item() + 1e-12
  zero = khamis_robust_loss(pred.detach(), torch.zeros_like(gt))                # no valid pixel: sum over nothing / max(0, 1)
  assert zero.item() == float(GOLD["khamis/loss_all_invalid"]) == 0.0


@pytest.mark.parametrize("shape,valid_frac", [((1, 1, 37, 90), 0.6), ((2, 1, 64, 130), 1.0), ((1, 1, 376, 1248), 0.3), ((1, 1, 16, 16), 0.0)])
def test_khamis_loss_matches_oracle(shape, valid_frac):
  g = torch.Generator().manual_seed(3)
  gt = torch.rand(shape, generator=g) * 60 + 0.5
  gt[torch.rand(shape, generator=g) >= valid_frac] = 0.0            # invalid pixels (gt == 0), incl. the all-invalid case
  pred = (gt + 4.0 * torch.randn(shape, generator=g)).abs()
  pr = pred.clone().requires_grad_(); pf = pred.clone().to(DEV).requires_grad_()
  if valid_frac > 0:
    ref = O.khamis_robust_loss(pr, gt)                              # the reference formulation (boolean indexing), on the CPU
    ref.backward()
    gref = pr.grad
  else:
    ref, gref = torch.zeros(()), torch.zeros_like(pr)               # sum over an empty mask / max(0, 1)
  out = khamis_robust_loss(pf, gt.to(DEV))
  (3.0 * out).backward()
  assert abs(out.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item())), (out.item(), ref.item())
  assert (pf.grad.cpu() / 3.0 - gref).abs().max().item() <= 1e-6 * max(gref.abs().max().item(), 1e-12) + 1e-12


@pytest.mark.parametrize("B,H,W", [(1, 37, 90), (3, 64, 130), (2, 376, 1248)])
def test_eval_metrics_match_oracle(B, H, W):
  g = torch.Generator().manual_seed(5)
  gt = torch.rand(B, 1, H, W, generator=g) * 60 + 0.5
  gt[torch.rand(B, 1, H, W, generator=g) < 0.4] = 0.0
  pred = gt + 3.0 * torch.randn(B, 1, H, W, generator=g)
  sums = ops.eval_metrics(pred.to(DEV), gt.to(DEV)).cpu().double()
  for b in range(B):
    ref = O.eval_metrics(pred[b:b + 1], gt[b:b + 1])                                             # train.py:98-107 restated
    nvalid = (gt[b] > 0).sum().item()
    assert sums[b, 1].item() == nvalid
    assert abs(sums[b, 0].item() / nvalid - ref["EPE"]) < 1e-5
    for i, t in enumerate(O.D1_THRESHOLDS):
      assert abs(sums[b, 2 + i].item() / nvalid - ref[f"D1_all_{t}px"]) < 1e-6


def test_photo_loss_per_sample_equals_one_pair_at_a_time():
  B, H, W = 3, 64, 130
  left, right, gt = O.make_stereo_pair(B, H, W, seed=1000, max_disp_px=30.0)
  disp = (gt.clamp(min=0) + 2.0 * torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(7)) + 1.0).to(DEV)
  l, r = left.to(DEV), right.to(DEV)
  loss, _ = ops.photo_loss(l, r, disp.squeeze(1).contiguous())
  for b in range(B):                                      # the oracle's restatement of adapt.py:78-86, one pair at a time, on the CPU
    ref = O.monodepth_single_loss(left[b:b + 1], right[b:b + 1], disp[b:b + 1].cpu())
    assert abs(loss[1 + b].item() - ref.item()) <= 3e-6 * max(1.0, abs(ref.item())), (b, loss[1 + b].item(), ref.item())
  ref_all = O.monodepth_single_loss(left, right, disp.cpu())
  assert abs(loss[0].item() - ref_all.item()) <= 3e-6


def test_validate_ovs_vs_reference_golden():
  """adapt.py:122-142: eval mode, no grad, one pair at a time in the reference -> batched here; the per-pair losses must be the
  ones the reference's own loop produced (gen_golden.py run_rows_case), and the networks must be back in train mode."""
  from stereonet_b200.validation import validate_ovs
  f, s = _nets()
  H, W, n = 96, 256, 5
  pairs = [O.make_stereo_pair(1, H, W, seed=2000 + i, max_disp_px=40.0)[:2] for i in range(n)]
  lefts = torch.cat([p[0] for p in pairs]).to(DEV); rights = torch.cat([p[1] for p in pairs]).to(DEV)
  f.train(); s.train()
  got = validate_ovs(f, s, lefts, rights, chunk=2).cpu().numpy()
  assert f.training and s.training
  ref = GOLD["validate/losses"]
  assert np.abs(got - ref).max() <= 2e-5, (got, ref)
  # running statistics untouched by validation
  sd = O.make_stereo_state(22, sharpen=10.0)
  assert torch.equal(s.state_dict()["filter.0.0.1.running_mean"].cpu(), sd["filter.0.0.1.running_mean"])


def test_evaluate_vs_reference_golden():
  """train.evaluate (train.py:74-126) over 3 batches of 2 pairs: EPE, D1-all at 2/3/4/5 px and FCS against the reference's values."""
  from stereonet_b200.validation import evaluate
  f, s = _nets()
  H, W = 96, 256
  batches = []
  for i in range(3):
    l, r, gt = O.make_stereo_pair(2, H, W, seed=3000 + i, max_disp_px=40.0)
    batches.append((l.to(DEV), r.to(DEV), gt.to(DEV)))
  got = evaluate(f, s, batches)
  assert f.training and s.training
  assert abs(got["EPE"] - float(GOLD["evaluate/EPE"])) < 1e-3                      # north star: |dEPE| <= 1e-3 px
  assert abs(got["FCS"] - float(GOLD["evaluate/FCS"])) < 1e-4 * max(1.0, abs(float(GOLD["evaluate/FCS"])))
  for i, t in enumerate((2, 3, 4, 5)):                                            # a pixel within 1e-2 px of a threshold may flip: 1 / 49152
    assert abs(got["D1_all_{}px".format(t)] - float(GOLD["evaluate/D1_all"][i])) < 1e-4


def test_replay_step_vs_reference_golden():
  """One experience-replay step exactly as adapt.py:328-349,381-394 runs it (two train-mode passes, Khamis loss on the replay
  sample, 1 : 0.05 mix, clip on stereo_net, Adam lr 5e-5) against the reference's own step: loss terms, clipped-gradient norm,
  BatchNorm running statistics after TWO updates, post-Adam weights."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  H, W = 96, 256
  f, s = _nets()
  l, r, _ = O.make_stereo_pair(1, H, W, seed=1000, max_disp_px=40.0)
  rl, rr, rgt = O.make_stereo_pair(1, H, W, seed=1001, max_disp_px=40.0)
  l, r, rl, rr, rgt = (t.to(DEV) for t in (l, r, rl, rr, rgt))
  opt = make_optimizer(f, s, lr=5e-5)
  st = AdaptStepper(f, s, opt, H, W, clip_grad_norm=False)
  f.train(); s.train()
  out = st.predict(l, r)
  l_mono = monodepth_single_loss(l, r, out, 0)
  out_er = st.predict(rl, rr)
  l_er = khamis_robust_loss(out_er["pred_disp_l/0"], rgt)
  assert abs(l_mono.item() - float(GOLD["er/loss_mono"])) < 2e-5
  assert abs(l_er.item() - float(GOLD["er/loss_replay"])) < 2e-4 * float(GOLD["er/loss_replay"])
  assert np.abs(out_er["pred_disp_l/0"].detach().cpu().numpy() - GOLD["er/pred_replay"]).max() <= 2e-2      # train-mode BN, 2nd pass
  opt.zero_grad()
  (l_mono + 0.05 * l_er).backward()
  gn = torch.nn.utils.clip_grad_norm_(s.parameters(), 1.0)
  assert abs(gn.item() - float(GOLD["er/grad_norm_stereo"])) <= 1.5e-2 * float(GOLD["er/grad_norm_stereo"])
  opt.step()
  for key in GOLD.files:
    if key.startswith("er/post/") and ("running_" in key or "num_batches" in key):
      _, _, tag, n = key.split("/", 3)
      v = (s if tag == "s" else f).state_dict()[n].cpu().numpy()
      if "num_batches" in key:
        # two train-mode passes per step; feature_net runs on the left and the right image of each; conv2 never runs
        assert int(v) == int(GOLD[key]) == (0 if ".conv2." in n else (2 if tag == "s" else 4)), key
      else:
        assert np.abs(v - GOLD[key]).max() <= 1e-4 * max(1.0, np.abs(GOLD[key]).max()), key
  d = np.abs(s.conv3d_alone.weight.detach().cpu().numpy() - GOLD["er/post/s/conv3d_alone.weight"])
  assert d.max() <= 2.1 * 5e-5 and (d <= 3e-6).mean() >= 0.97, (d.max(), (d <= 3e-6).mean())


def test_batched_replay_step():
  """Row f3: the stream frame and the replay sample share one batch-2 pass.  The loss must be Monodepth(frame) +
  0.05 * Khamis(replay) of that pass's own predictions, every used parameter must receive a gradient, and the update must
  match a step whose two loss terms are evaluated by the PyTorch formulas on the same batched forward."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  H, W = 96, 256
  l, r, _ = O.make_stereo_pair(1, H, W, seed=1000, max_disp_px=40.0)
  rl, rr, rgt = O.make_stereo_pair(1, H, W, seed=1001, max_disp_px=40.0)
  l, r, rl, rr, rgt = (t.to(DEV) for t in (l, r, rl, rr, rgt))
  res = []
  for fused in (True, False):
    f, s = _nets()
    opt = make_optimizer(f, s, lr=5e-5)
    st = AdaptStepper(f, s, opt, H, W, batched_replay=True)
    if fused:
      loss, fcs, out = st.step(l, r, replay=(rl, rr, rgt))
    else:   # the same batched pass with the two loss terms evaluated by the ORACLE's formulas (torch autograd on the GPU tensors)
      f.train(); s.train()
      o = st.predict(torch.cat([l, rl]), torch.cat([r, rr]))
      loss = O.monodepth_single_loss(l, r, o["pred_disp_l/0"][:1]) + 0.05 * O.khamis_robust_loss(o["pred_disp_l/0"][1:], rgt)
      opt.zero_grad(); loss.backward(); st._update()
    torch.cuda.synchronize()
    res.append((loss.item(), {n: p.detach().cpu().clone() for n, p in s.named_parameters()}))
    if fused:
      assert tuple(out["pred_disp_l/0"].shape) == (1, 1, H, W) and torch.isfinite(fcs).item()
      missing = [n for n, p in list(s.named_parameters()) + list(f.named_parameters()) if p.grad is None and ".conv2." not in n]
      assert not missing, missing
  (la, wa), (lb, wb) = res
  assert abs(la - lb) <= 2e-5 * max(1.0, abs(lb)), (la, lb)
  for n in wa:
    assert (wa[n] - wb[n]).abs().max().item() <= 2.1 * 5e-5, n      # one Adam step: entries differ by at most 2 * lr


def test_replay_step_cuda_graph_matches_eager():
  """The ER step (adapt.py:339-349: second pass on a replay sample, Khamis loss, 1 : 0.05 mix) replayed as a CUDA graph
  must leave the same weights as the eager step with the reference's boolean-index loss, on changing frames / replay samples."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  H, W = 96, 256
  frames = [O.make_stereo_pair(1, H, W, seed=1000 + i, max_disp_px=40.0) for i in range(3)]
  replays = [O.make_stereo_pair(1, H, W, seed=2000 + i, max_disp_px=40.0) for i in range(3)]
  res = []
  for use_graph in (False, True):
    f, s = _nets()
    st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, capturable=True), H, W, use_graph=use_graph)
    losses = []
    for (l, r, _), (rl, rr, rgt) in zip(frames, replays):
      loss, _, _ = st.step(l.to(DEV), r.to(DEV), replay=(rl.to(DEV), rr.to(DEV), rgt.to(DEV)))
      losses.append(loss.item())
    torch.cuda.synchronize()
    res.append((losses, {n: v.detach().cpu().clone() for n, v in list(s.state_dict().items()) + list(f.state_dict().items())}))
  (le, we), (lg, wg) = res
  # step 1 runs on identical weights; later steps on weights that have drifted by up to 2*lr per entry (Adam's first steps move
  # every weight by ~lr * sign(g), and rounding-level gradient differences flip signs of near-zero gradients)
  assert abs(le[0] - lg[0]) < 2e-5 and max(abs(a - b) for a, b in zip(le, lg)) < 3e-4, (le, lg)
  for n in we:
    if "num_batches_tracked" in n:
      assert torch.equal(we[n], wg[n]), n
    else:
      assert (we[n] - wg[n]).abs().max().item() <= 3 * 2.1 * 5e-5 + 1e-5 * we[n].abs().max().item(), n
