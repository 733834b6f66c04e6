"""GPU parity of SURVEY.md section 8 rows f3 / f4: the Khamis replay loss, the batched replay step, the device-resident
evaluation metrics and the batched OVS validation, against the plain-PyTorch restatement of the reference formulas
(loss_functions.py:6-15, train.py:98-107, adapt.py:122-142,339-349)."""
import pytest
import torch

import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import ops
from stereonet_b200.losses import (LinearWarping, khamis_robust_loss, khamis_robust_loss_fused, monodepth_single_loss,
                                   feature_contrast_mean)
from test_gpu_kernels import DEV

pytestmark = pytest.mark.gpu


def _nets(k=3, sharpen=10.0):
  f = S.FeatureExtractorNetwork(k).to(DEV); s = S.StereoNet(k, 1, 0).to(DEV)
  f.load_state_dict(O.make_feature_state(k, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=sharpen))
  return f, s


@pytest.mark.parametrize("shape,valid_frac", [((1, 1, 37, 90), 0.6), ((2, 1, 64, 130), 1.0), ((1, 1, 376, 1248), 0.3), ((1, 1, 16, 16), 0.0)])
def test_khamis_loss_matches_reference_formula(shape, valid_frac):
  g = torch.Generator().manual_seed(3)
  gt = torch.rand(shape, generator=g) * 60 + 0.5
  gt[torch.rand(shape, generator=g) >= valid_frac] = 0.0            # invalid pixels (gt == 0), incl. the all-invalid case
  pred = (gt + 4.0 * torch.randn(shape, generator=g)).abs()
  pr = pred.clone().to(DEV).requires_grad_(); pf = pred.clone().to(DEV).requires_grad_()
  gtd = gt.to(DEV)
  if valid_frac > 0:
    ref = khamis_robust_loss(pr, gtd)                               # the reference's boolean-index formulation
    ref.backward()
    gref = pr.grad
  else:
    ref, gref = torch.zeros((), device=DEV), torch.zeros_like(pr)    # sum over an empty mask / max(0, 1)
  out = khamis_robust_loss_fused(pf, gtd)
  (3.0 * out).backward()
  assert abs(out.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item())), (out.item(), ref.item())
  assert (pf.grad / 3.0 - gref).abs().max().item() <= 1e-6 * max(gref.abs().max().item(), 1e-12) + 1e-12


@pytest.mark.parametrize("B,H,W", [(1, 37, 90), (3, 64, 130), (2, 376, 1248)])
def test_eval_metrics_match_train_evaluate_formulas(B, H, W):
  g = torch.Generator().manual_seed(5)
  gt = torch.rand(B, 1, H, W, generator=g) * 60 + 0.5
  gt[torch.rand(B, 1, H, W, generator=g) < 0.4] = 0.0
  pred = gt + 3.0 * torch.randn(B, 1, H, W, generator=g)
  sums = ops.eval_metrics(pred.to(DEV), gt.to(DEV)).cpu().double()
  for b in range(B):
    valid = gt[b] > 0
    err = (pred[b] - gt[b]).abs()
    assert sums[b, 1].item() == valid.sum().item()
    assert abs(sums[b, 0].item() / sums[b, 1].item() - err[valid].mean().item()) < 1e-5          # EPE, train.py:103
    for i, t in enumerate((2, 3, 4, 5)):                                                          # D1-all, train.py:106-107
      assert sums[b, 2 + i].item() == (valid * (err > t)).sum().item()


def test_photo_loss_per_sample_equals_one_pair_at_a_time():
  B, H, W = 3, 64, 130
  left, right, gt = O.make_stereo_pair(B, H, W, seed=1000, max_disp_px=30.0)
  disp = (gt.clamp(min=0) + 2.0 * torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(7)) + 1.0).to(DEV)
  l, r = left.to(DEV), right.to(DEV)
  loss, _ = ops.photo_loss(l, r, disp.squeeze(1).contiguous())
  warper = LinearWarping(H, W, torch.device(DEV))
  for b in range(B):
    ref = monodepth_single_loss(l[b:b + 1], r[b:b + 1], {"pred_disp_l/0": disp[b:b + 1]}, warper, 0)
    assert abs(loss[1 + b].item() - ref.item()) <= 3e-6 * max(1.0, abs(ref.item())), (b, loss[1 + b].item(), ref.item())
  ref_all = monodepth_single_loss(l, r, {"pred_disp_l/0": disp}, warper, 0)
  assert abs(loss[0].item() - ref_all.item()) <= 3e-6


def test_validate_ovs_matches_reference_loop():
  """adapt.py:122-142: eval mode, no grad, one pair at a time -> here batched; same per-pair losses, networks back in train mode."""
  from stereonet_b200.validation import validate_ovs
  f, s = _nets()
  H, W, n = 96, 256, 5
  pairs = [O.make_stereo_pair(1, H, W, seed=2000 + i, max_disp_px=40.0)[:2] for i in range(n)]
  lefts = torch.cat([p[0] for p in pairs]).to(DEV); rights = torch.cat([p[1] for p in pairs]).to(DEV)
  f.train(); s.train()
  got = validate_ovs(f, s, lefts, rights, chunk=2)
  assert f.training and s.training
  warper = LinearWarping(H, W, torch.device(DEV))
  f.eval(); s.eval()
  with torch.no_grad():
    for i in range(n):
      l, r = lefts[i:i + 1], rights[i:i + 1]
      out = s(l, f(l), f(r), "l", output_cost_volume=True)
      ref = monodepth_single_loss(l, r, out, warper, 0)
      assert abs(got[i].item() - ref.item()) <= 5e-6 * max(1.0, abs(ref.item())), (i, got[i].item(), ref.item())


def test_evaluate_matches_train_evaluate():
  from stereonet_b200.validation import evaluate
  f, s = _nets()
  H, W = 96, 256
  batches = []
  for i in range(3):
    l, r, gt = O.make_stereo_pair(2, H, W, seed=3000 + i, max_disp_px=40.0)
    batches.append((l.to(DEV), r.to(DEV), gt.to(DEV)))
  got = evaluate(f, s, batches)
  assert f.training and s.training
  # the reference loop (train.py:89-110) on the same model outputs
  f.eval(); s.eval()
  epe, d1, fcs = [], [], []
  with torch.no_grad():
    for l, r, gt in batches:
      out = s(l, f(l), f(r), "l", output_cost_volume=True)
      pred = out["pred_disp_l/0"]
      valid = gt > 0
      epe.append(torch.abs(pred - gt)[valid].mean().item())
      d1.append([((valid * (torch.abs(pred - gt) > t)).sum() / float(valid.sum())).item() for t in (2, 3, 4, 5)])
      cv = out["cost_volume_l/3"]
      srt = torch.sort(cv, dim=1, descending=True)[0]
      fcs.append((srt[:, 0] - srt[:, 2:].mean(dim=1)).mean().item())
  assert abs(got["EPE"] - sum(epe) / 3) < 1e-4
  assert abs(got["FCS"] - sum(fcs) / 3) < 1e-5 * max(1.0, abs(sum(fcs) / 3))
  for i, t in enumerate((2, 3, 4, 5)):
    assert abs(got["D1_all_{}px".format(t)] - sum(x[i] for x in d1) / 3) < 1e-6


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
    st = AdaptStepper(f, s, opt, H, W, fused_loss=fused, batched_replay=True)
    if fused:
      loss, fcs, out = st.step(l, r, replay=(rl, rr, rgt))
    else:   # the same batched pass with the reference formulas (boolean-index Khamis loss, PyTorch Monodepth loss)
      f.train(); s.train()
      o = st.predict(torch.cat([l, rl]), torch.cat([r, rr]))
      head = {k: v[:1] for k, v in o.items()}
      loss = monodepth_single_loss(l, r, head, st.warper, 0) + 0.05 * khamis_robust_loss(o["pred_disp_l/0"][1:], rgt)
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
    st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, capturable=True), H, W, use_graph=use_graph, fused_loss=True)
    losses = []
    for (l, r, _), (rl, rr, rgt) in zip(frames, replays):
      loss, _, _ = st.step(l.to(DEV), r.to(DEV), replay=(rl.to(DEV), rr.to(DEV), rgt.to(DEV)))
      losses.append(loss.item())
    torch.cuda.synchronize()
    res.append((losses, {n: v.detach().cpu().clone() for n, v in list(s.state_dict().items()) + list(f.state_dict().items())}))
  (le, we), (lg, wg) = res
  assert max(abs(a - b) for a, b in zip(le, lg)) < 2e-5, (le, lg)
  for n in we:
    if "num_batches_tracked" in n:
      assert torch.equal(we[n], wg[n]), n
    else:
      assert (we[n] - wg[n]).abs().max().item() <= 3 * 2.1 * 5e-5 + 1e-5 * we[n].abs().max().item(), n
