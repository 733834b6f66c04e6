"""GPU parity of the backward kernels / autograd Functions against torch autograd on the CPU (the oracle), layer by
layer and end to end against the adaptation-step golden vectors produced by the reference (tests/golden)."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200.autograd import functions as Fn, fused
from test_gpu_kernels import cl, uncl, rnd, close, DEV
from test_oracle_golden import CASES, TRAIN_SHARPEN_RATIO, build, summ

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def randomize_bn(bn, seed):
  with torch.no_grad():
    bn.weight.copy_(rnd(32, seed=seed).abs() + 0.5); bn.bias.copy_(rnd(32, seed=seed + 1))
    bn.running_mean.copy_(rnd(32, seed=seed + 2) * 0.1); bn.running_var.copy_(rnd(32, seed=seed + 3).abs() + 0.5)


@pytest.mark.parametrize("backend", ["ffma", "tc3"])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("three_d,dil,residual", [(False, 1, True), (False, 4, True), (True, 1, False)])
def test_conv_bn_lrelu_backward(three_d, dil, residual, training, backend):
  shape = (1, 32, 6, 9, 21) if three_d else (2, 32, 17, 29)
  # Small-integer inputs and dyadic weights make every convolution sum exact in fp32 (and in the 3xTF32 split), so z is
  # bit-identical on CPU and GPU and no activation sits within rounding distance of the LeakyReLU kink (a flipped
  # mask there is a legitimate O(1) difference between any two fp32 implementations, not an error).
  x = torch.randint(-2, 3, shape, generator=torch.Generator().manual_seed(1)).float().requires_grad_()
  torch.manual_seed(5)
  conv = (torch.nn.Conv3d(32, 32, 3, padding=1) if three_d else torch.nn.Conv2d(32, 32, 3, padding=dil, dilation=dil))
  with torch.no_grad():
    conv.weight.copy_(torch.randint(-2, 3, conv.weight.shape, generator=torch.Generator().manual_seed(3)).float() / 16)
    conv.bias.copy_(torch.randint(-4, 5, (32,), generator=torch.Generator().manual_seed(4)).float() / 8)
  bn = torch.nn.BatchNorm3d(32) if three_d else torch.nn.BatchNorm2d(32)
  randomize_bn(bn, 7)
  gconv, gbn = copy.deepcopy(conv).to(DEV), copy.deepcopy(bn).to(DEV)
  conv.train(training); bn.train(training); gbn.train(training)
  y = F.leaky_relu(bn(conv(x)), 0.2)
  if residual:
    y = x + y
  gy = rnd(*shape, seed=2)
  y.backward(gy)
  old = fused.CONV_BACKEND
  fused.set_conv_backend(backend)
  try:
    xg = cl(x.detach()).requires_grad_()
    yg = Fn.conv_bn_lrelu_autograd(xg, gconv, gbn, dil, residual, training)
    yg.backward(cl(gy))
  finally:
    fused.set_conv_backend(old)
  close(uncl(yg.detach()), y.detach(), 1e-5, "forward")
  close(uncl(xg.grad), x.grad, 2e-5, "dx")
  close(gconv.weight.grad.cpu(), conv.weight.grad, 2e-5, "dW")
  close(gbn.weight.grad.cpu(), bn.weight.grad, 2e-5, "dgamma")
  close(gbn.bias.grad.cpu(), bn.bias.grad, 2e-5, "dbeta")
  if training:   # the conv-bias gradient under batch-stat BN is mathematically zero: compare absolutely
    assert gconv.bias.grad.abs().max().item() <= 1e-3 * (conv.weight.grad.abs().max().item() + 1.0)
  else:
    close(gconv.bias.grad.cpu(), conv.bias.grad, 2e-5, "dbias")


@pytest.mark.parametrize("ksize,stride,H,W", [(5, 2, 23, 37), (5, 2, 24, 40), (3, 1, 12, 19)])
def test_conv_plain_backward(ksize, stride, H, W):
  x = rnd(2, 32, H, W, seed=1).requires_grad_()
  torch.manual_seed(5)
  conv = torch.nn.Conv2d(32, 32, ksize, stride=stride, padding=ksize // 2)
  gconv = copy.deepcopy(conv).to(DEV)
  y = conv(x)
  gy = rnd(*y.shape, seed=2)
  y.backward(gy)
  xg = cl(x.detach()).requires_grad_()
  yg = Fn.ConvC32.apply(xg, gconv.weight, gconv.bias, gconv, ksize, stride)
  yg.backward(cl(gy))
  close(uncl(yg.detach()), y.detach(), 1e-5, "forward")
  close(uncl(xg.grad), x.grad, 2e-5, "dx")
  close(gconv.weight.grad.cpu(), conv.weight.grad, 2e-5, "dW")
  close(gconv.bias.grad.cpu(), conv.bias.grad, 2e-5, "dbias")


def test_first_conv_backward():
  img = torch.rand(2, 3, 37, 75, generator=torch.Generator().manual_seed(1))
  torch.manual_seed(5)
  conv = torch.nn.Conv2d(3, 32, 5, stride=2, padding=2)
  gconv = copy.deepcopy(conv).to(DEV)
  y = conv(img)
  gy = rnd(*y.shape, seed=2)
  y.backward(gy)
  yg = Fn.Conv5x5s2First.apply(img.to(DEV), gconv.weight, gconv.bias)
  yg.backward(cl(gy))
  close(gconv.weight.grad.cpu(), conv.weight.grad, 2e-5, "dW")
  close(gconv.bias.grad.cpu(), conv.bias.grad, 2e-5, "dbias")


def test_head_backward():
  B, D, H, W = 2, 12, 7, 33
  x = rnd(B, 32, D, H, W, seed=1).requires_grad_()
  torch.manual_seed(5)
  conv = torch.nn.Conv3d(32, 1, 3, padding=1)
  with torch.no_grad():
    conv.weight.mul_(10.0)
  gconv = copy.deepcopy(conv).to(DEV)
  cost = conv(x).squeeze(1)
  pred = O.soft_argmin(cost)
  gp, gc = rnd(B, H, W, seed=2), rnd(B, D, H, W, seed=3) * 0.1
  ((pred * gp).sum() + (cost * gc).sum()).backward()
  xg = cl(x.detach()).requires_grad_()
  cg, pg, _fcs = Fn.Conv3dOutSoftargmin.apply(xg, gconv.weight, gconv.bias)
  ((pg * gp.to(DEV)).sum() + (cg * gc.to(DEV)).sum()).backward()
  close(uncl(xg.grad), x.grad, 3e-5, "dx")
  close(gconv.weight.grad.cpu(), conv.weight.grad, 3e-5, "dw")
  close(gconv.bias.grad.cpu(), conv.bias.grad, 3e-5, "db")


def test_cost_volume_and_upsample_autograd():
  L, R = rnd(1, 32, 5, 37, seed=1).requires_grad_(), rnd(1, 32, 5, 37, seed=2).requires_grad_()
  gc = rnd(1, 32, 24, 5, 37, seed=3)
  O.cost_volume(L, R, 24).backward(gc)
  Lg, Rg = cl(L.detach()).requires_grad_(), cl(R.detach()).requires_grad_()
  Fn.CostVolume.apply(Lg, Rg, 24).backward(cl(gc))
  close(uncl(Lg.grad), L.grad, 1e-6, "dL"); close(uncl(Rg.grad), R.grad, 1e-6, "dR")
  p = rnd(2, 9, 15, seed=4).requires_grad_()
  g = rnd(2, 68, 120, seed=5)
  (8.0 * F.interpolate(p.unsqueeze(1), size=(68, 120), mode="bilinear", align_corners=False).squeeze(1)).backward(g)
  pg = p.detach().to(DEV).requires_grad_()
  Fn.Upsample.apply(pg, 68, 120, 8.0).backward(g.to(DEV))
  close(pg.grad.cpu(), p.grad, 1e-5, "dpred")


@pytest.mark.parametrize("training", [True, False])
def test_refinement_backward(training):
  B, h, w, H, W = 1, 9, 15, 68, 120
  sd = O.clone_state(O.make_stereo_state(22, 10.0), True)
  coarse = (torch.rand(B, h, w, generator=torch.Generator().manual_seed(1)) * 20).requires_grad_()
  rgb = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(2))
  out = O.refinement(sd, coarse, rgb, training)
  g = rnd(B, 1, H, W, seed=3)
  out.backward(g)
  net = S.StereoNet(3, 1, 0).to(DEV)
  net.load_state_dict({k: v.detach() for k, v in O.make_stereo_state(22, 10.0).items()})
  net.train(training)
  ref = net.edge_aware_refinements[0]
  cg = coarse.detach().to(DEV).requires_grad_()
  og = ref(cg, rgb.to(DEV))
  og.backward(g.to(DEV))
  close(og.detach().cpu(), out.detach(), 1e-5, "refinement forward")
  close(cg.grad.cpu(), coarse.grad, 5e-3, "dcoarse")
  p = "edge_aware_refinements.0."
  for n, prm in ref.named_parameters():
    if ".conv2." in n:
      assert prm.grad is None
      continue
    if training and n.endswith("0.0.bias"):
      continue                               # conv bias under batch-stat BN: zero up to rounding in both
    # random-float activations: a handful of LeakyReLU masks within rounding distance of the kink may flip between the
    # two fp32 implementations (each flip moves a reduced gradient by ~1/sqrt(N)), hence the looser bound here; the
    # exact-arithmetic layer tests above pin the kernels to 2e-5.
    close(prm.grad.cpu(), sd[p + n].grad, 2e-2 if training else 5e-3, n)


@pytest.mark.parametrize("loss_impl", ["oracle", "fused"])
@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["train"]])
def test_adapt_step_vs_reference_golden(name, loss_impl):
  """One full adaptation step (adapt.py:313-337,381-394) on the GPU vs the reference's own step (golden vectors).

  Tolerances: the reference gradient itself is ill-conditioned at this size — perturbing the input images by 1e-7
  (relative, below one ulp of most pixels) moves its own filter.* / conv3d_alone gradients by cos 0.9986-0.9993, 3-4 %
  on single entries and up to 1 % in norm (LeakyReLU masks within rounding distance of the kink flip under train-mode
  BN; measured with the oracle on the CPU).  The bounds below sit just outside that band; the
  layer-level tests in this file, which use exactly representable data, pin every backward kernel to 2e-5."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  from stereonet_b200.losses import monodepth_single_loss
  cfg = CASES[name]
  g = np.load(os.path.join(GOLD, name + ".npz"))
  fsd, _, left, right, _ = build(cfg)
  ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"] * TRAIN_SHARPEN_RATIO)
  f = S.FeatureExtractorNetwork(cfg["k"]).to(DEV); s = S.StereoNet(cfg["k"], 1, cfg["s"]).to(DEV)
  f.load_state_dict(fsd); s.load_state_dict(ssd)
  opt = make_optimizer(f, s, lr=5e-5)
  stepper = AdaptStepper(f, s, opt, cfg["H"], cfg["W"], clip_grad_norm=False)
  f.train(); s.train()
  l, r = left.to(DEV), right.to(DEV)
  outputs = stepper.predict(l, r)
  if loss_impl == "fused":     # snb_photo_loss (row f1, the product path) against the reference's own loss value and gradients
    loss = monodepth_single_loss(l, r, outputs, cfg["s"])
  else:                        # the oracle's torch restatement of adapt.py:78-86 seeds the CUDA backward instead (isolates the model)
    loss = O.monodepth_single_loss(l, r, outputs["pred_disp_l/{}".format(cfg["s"])])
  opt.zero_grad()
  loss.backward()
  print(f"[parity] {name} loss {loss.item():.7f} vs reference {float(g['train/loss']):.7f}")
  assert abs(loss.item() - float(g["train/loss"])) < 2e-5
  worst = 0.0
  for tag, net in (("s", s), ("f", f)):
    for n, p in net.named_parameters():
      if f"grad_none/{tag}/{n}" in g.files:
        assert p.grad is None, n
        continue
      if n.endswith(".0.0.bias") or n == "conv3d_alone.bias":
        continue        # conv bias feeding a batch-stat BN, and the bias under a softmax (shift invariant): mathematically
                        # zero gradients, pure rounding noise in both implementations
      ref = g[f"grad_sum/{tag}/{n}"]
      got = summ(p.grad.cpu())
      if ref[2] < 1e-6:
        assert got[2] < 1e-5, n          # structurally ~zero gradient (e.g. conv_alone.bias cancels in L - R): noise only
        continue
      rel = abs(got[2] - ref[2]) / (ref[2] + 1e-12)
      worst = max(worst, rel)
      # one-element tensors (conv2d_out.bias: a cancelling sum over all pixels) are "single entries": same 6 % band as the
      # element-wise bound below; multi-element tensors are held to 3 % in norm
      assert rel <= (6e-2 if p.numel() == 1 else 3e-2), (n, got, ref)
      key = f"grad/{tag}/{n}"
      if key in g.files:       # full tensors: direction and magnitude (LeakyReLU kink flips allow ~1 % on single entries)
        a, b = p.grad.cpu().numpy().ravel().astype(np.float64), g[key].ravel().astype(np.float64)
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        assert cos >= 0.997, (key, cos)
        # Single entries: measured on the B200 (scripts/grad_sensitivity.py, round 2) the three fp32-grade convolution back
        # ends of this library (fp16 split, 3xTF32, plain FFMA) differ from EACH OTHER by up to 6 % on single entries of
        # k3_b2_sharp and from the reference by 10.4 % / 8.6 % / 9.3 % (k3_ragged: 1.0-1.4 %): the band is a property of the
        # case (LeakyReLU masks at rounding distance from the kink under batch-stat BN), not of a kernel.
        assert np.abs(a - b).max() <= 1.2e-1 * (np.abs(b).max() + 1e-12), key
  print(f"[parity] {name} worst relative grad-norm error {worst:.3e}")
  gn = torch.nn.utils.clip_grad_norm_(s.parameters(), 1.0)
  assert abs(gn.item() - float(g["train/grad_norm_stereo"])) <= 1.5e-2 * float(g["train/grad_norm_stereo"])
  opt.step()
  for key in g.files:
    if key.startswith("post/") and "running_" not in key and "num_batches" not in key and not key.endswith(".bias"):
      # (conv_alone.bias cancels in L - R: its gradient is rounding noise, which Adam's first step turns into +-lr)
      _, tag, n = key.split("/", 2)
      sd = (s if tag == "s" else f).state_dict()
      # first Adam step moves every weight by lr * g/(|g| + eps): entries whose tiny gradient changes sign differ by 2*lr
      d = np.abs(sd[n].cpu().numpy() - g[key])
      assert d.max() <= 2.1 * 5e-5 and (d <= 3e-6).mean() >= 0.97, (key, d.max(), (d <= 3e-6).mean())


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("bn_bias", [60.0, -60.0])
def test_refine_head_backward_exact_side_of_kink(training, bn_bias):
  """RefineHead (upsample + concat + conv 4->32 + BN + LeakyReLU) with the BN shift pushed far from the LeakyReLU kink, so
  every mask is unambiguous and gradients must agree to fp32 rounding."""
  B, h, w, H, W = 2, 9, 15, 68, 120
  coarse = (torch.rand(B, h, w, generator=torch.Generator().manual_seed(1)) * 20).requires_grad_()
  rgb = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(2))
  torch.manual_seed(5)
  conv = torch.nn.Conv2d(4, 32, 3, padding=1); bn = torch.nn.BatchNorm2d(32)
  randomize_bn(bn, 7)
  with torch.no_grad():
    bn.bias.fill_(bn_bias)
    if not training:
      bn.running_var.fill_(400.0)          # keep |BN(z)| << |shift| with running statistics as well
  gconv, gbn = copy.deepcopy(conv).to(DEV), copy.deepcopy(bn).to(DEV)
  bn.train(training); gbn.train(training)
  up = F.interpolate(coarse.unsqueeze(1), size=(H, W), mode="bilinear", align_corners=False) * (W / w)
  y = F.leaky_relu(bn(conv(torch.cat([up, rgb], 1))), 0.2)
  gy, gu = rnd(B, 32, H, W, seed=3), rnd(B, H, W, seed=4)
  ((y * gy).sum() + (up.squeeze(1) * gu).sum()).backward()
  cg = coarse.detach().to(DEV).requires_grad_()
  upg, yg = Fn.refine_head_autograd(cg, rgb.to(DEV), gconv, gbn, training)
  ((yg * cl(gy)).sum() + (upg * gu.to(DEV)).sum()).backward()
  close(uncl(yg.detach()), y.detach(), 1e-5, "forward")
  close(cg.grad.cpu(), coarse.grad, 3e-5, "dcoarse")
  close(gconv.weight.grad.cpu(), conv.weight.grad, 3e-5, "dW")
  close(gbn.weight.grad.cpu(), bn.weight.grad, 3e-5, "dgamma")
  close(gbn.bias.grad.cpu(), bn.bias.grad, 3e-5, "dbeta")


def test_refine_tail_backward():
  B, H, W = 2, 23, 41
  x = rnd(B, 32, H, W, seed=1).requires_grad_()
  up = (rnd(B, H, W, seed=2) * 2).requires_grad_()
  torch.manual_seed(5)
  conv = torch.nn.Conv2d(32, 1, 3, padding=1)
  gconv = copy.deepcopy(conv).to(DEV)
  out = F.relu(up.unsqueeze(1) + conv(x)).squeeze(1)
  g = rnd(B, H, W, seed=3)
  out.backward(g)
  xg, ug = cl(x.detach()).requires_grad_(), up.detach().to(DEV).requires_grad_()
  og = Fn.refine_tail_autograd(xg, gconv, ug)
  og.backward(g.to(DEV))
  close(og.detach().cpu(), out.detach(), 1e-5, "forward")
  keep = (out.detach().abs() > 1e-4) | (out.detach() == 0)        # ignore pixels sitting on the ReLU kink
  assert keep.float().mean() > 0.999
  close(ug.grad.cpu() * keep, up.grad * keep, 1e-5, "dup")
  close(uncl(xg.grad), x.grad, 1e-4, "dx")
  close(gconv.weight.grad.cpu(), conv.weight.grad, 1e-4, "dw")
  close(gconv.bias.grad.cpu(), conv.bias.grad, 1e-4, "db")


def test_adapt_step_cuda_graph_matches_eager():
  """AdaptStepper(use_graph=True) replays the whole update (fwd, loss, bwd, clip, Adam) as one CUDA graph: three steps
  on changing frames must leave the same weights / BN statistics / Adam state as three eager steps, and capturing must
  not advance the adaptation."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  k, Hh, Ww = 3, 96, 256
  fsd, ssd = O.make_feature_state(k, 11), O.make_stereo_state(22, sharpen=10.0)
  frames = [O.make_stereo_pair(1, Hh, Ww, seed=1000 + i, max_disp_px=40.0)[:2] for i in range(3)]
  results = []
  for use_graph in (False, True):
    f = S.FeatureExtractorNetwork(k).to(DEV); s = S.StereoNet(k, 1, 0).to(DEV)
    f.load_state_dict(fsd); s.load_state_dict(ssd)
    opt = make_optimizer(f, s, lr=5e-5, capturable=True)
    st = AdaptStepper(f, s, opt, Hh, Ww, clip_grad_norm=True, use_graph=use_graph)
    losses = []
    for l, r in frames:
      loss, fcs, out = st.step(l.to(DEV), r.to(DEV))
      losses.append(loss.item())
    torch.cuda.synchronize()
    # an eval-mode forward after the replays must see the updated weights (derived-weight caches invalidated)
    f.eval(); s.eval()
    with torch.no_grad():
      l, r = frames[0][0].to(DEV), frames[0][1].to(DEV)
      disp = s(l, f(l), f(r), "l")["pred_disp_l/0"].cpu()
    results.append((losses, {n: v.detach().cpu().clone() for n, v in list(s.state_dict().items()) + list(f.state_dict().items())}, disp))
  (le, we, de), (lg, wg, dg) = results
  assert max(abs(a - b) for a, b in zip(le, lg)) < 1e-5, (le, lg)
  for n in we:
    if "num_batches_tracked" in n:
      assert torch.equal(we[n], wg[n]), n
    else:
      assert (we[n] - wg[n]).abs().max().item() <= 2e-6 + 1e-5 * we[n].abs().max().item(), n
  assert (de - dg).abs().max().item() < 1e-2


@pytest.mark.parametrize("B,H,W", [(1, 37, 90), (2, 64, 130), (1, 376, 1248)])
def test_fused_photo_loss_matches_oracle(B, H, W):
  """snb_photo_loss (warp + SSIM + L1 + edge-aware smoothness + masked mean, value and d/d disp fused)
  against the oracle's plain-PyTorch restatement of the reference loss (O.monodepth_single_loss, autograd)."""
  from stereonet_b200.losses import monodepth_single_loss
  left, right, gt = O.make_stereo_pair(B, H, W, seed=1000, max_disp_px=min(60.0, W / 4))
  g = torch.Generator().manual_seed(5)
  disp = (gt.clamp(min=0) + 3.0 * torch.rand(B, 1, H, W, generator=g) + 1.0).to(DEV)       # positive, textured, some pixels invalid
  l, r = left.to(DEV), right.to(DEV)
  d_ref = disp.clone().requires_grad_()
  loss_ref = O.monodepth_single_loss(l, r, d_ref)
  loss_ref.backward()
  d_fused = disp.clone().requires_grad_()
  loss_fused = monodepth_single_loss(l, r, {"pred_disp_l/0": d_fused}, 0)
  (2.0 * loss_fused).backward()                                       # exercises the grad_output scaling
  assert abs(loss_fused.item() - loss_ref.item()) <= 2e-6 * max(1.0, abs(loss_ref.item())), (loss_fused.item(), loss_ref.item())
  gr, gf = d_ref.grad, d_fused.grad / 2.0
  scale = gr.abs().max().item()
  # |x| and clamp kinks: a pixel whose L - I^ (or SSIM clamp argument) sits within rounding distance of 0 may take the other
  # sub-gradient in either implementation; everything else must agree to fp32 rounding
  diff = (gr - gf).abs()
  assert (diff > 1e-4 * scale).float().mean().item() < 1e-3, (diff.max().item(), scale)
  cos = torch.nn.functional.cosine_similarity(gr.flatten(), gf.flatten(), dim=0).item()
  assert cos > 0.99999, cos


def test_adapt_step_fused_loss_matches_oracle_loss():
  """Three graph-replayed adaptation steps (snb_photo_loss inside) leave the same weights as three eager steps whose loss is
  the oracle's plain-PyTorch restatement of adapt.py:78-86 (autograd seeds the same CUDA backward)."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  k, Hh, Ww = 3, 96, 256
  fsd, ssd = O.make_feature_state(k, 11), O.make_stereo_state(22, sharpen=10.0)
  frames = [O.make_stereo_pair(1, Hh, Ww, seed=1000 + i, max_disp_px=40.0)[:2] for i in range(3)]
  res = []
  for fused_loss in (False, True):
    f = S.FeatureExtractorNetwork(k).to(DEV); s = S.StereoNet(k, 1, 0).to(DEV)
    f.load_state_dict(fsd); s.load_state_dict(ssd)
    opt = make_optimizer(f, s, lr=5e-5, capturable=True)
    st = AdaptStepper(f, s, opt, Hh, Ww, use_graph=fused_loss)
    if fused_loss:
      losses = [st.step(l.to(DEV), r.to(DEV))[0].item() for l, r in frames]
    else:
      losses = []
      f.train(); s.train()
      for l, r in frames:
        l, r = l.to(DEV), r.to(DEV)
        loss = O.monodepth_single_loss(l, r, st.predict(l, r)["pred_disp_l/0"])
        opt.zero_grad(); loss.backward(); st._update()
        losses.append(loss.item())
    torch.cuda.synchronize()
    res.append((losses, {n: v.detach().cpu().clone() for n, v in s.state_dict().items()}))
  (la, wa), (lb, wb) = res
  # step 1 runs on identical weights; later steps on weights that may have drifted by up to 2*lr per entry (below)
  assert abs(la[0] - lb[0]) < 2e-5 and max(abs(a - b) for a, b in zip(la, lb)) < 3e-4, (la, lb)
  # Adam's first steps move every weight by ~lr * sign(gradient): entries whose tiny gradient differs in rounding (the
  # backward pass is ill-conditioned through the LeakyReLU / batch-stat BN kinks) land up to 2*lr apart,
  # so compare the accumulated UPDATE as a direction and bound single entries by the step budget (3 steps x 2 x lr).
  ua, ub = [], []
  for n in wa:
    if "num_batches_tracked" in n or "running_" in n:
      continue
    d = (wa[n] - wb[n]).abs()
    assert d.max().item() <= 3 * 2.1 * 5e-5, (n, d.max().item())
    ua.append((wa[n] - ssd[n]).flatten()); ub.append((wb[n] - ssd[n]).flatten())
  ua, ub = torch.cat(ua).double(), torch.cat(ub).double()
  cos = float(ua @ ub / (ua.norm() * ub.norm() + 1e-30))
  assert cos > 0.97, cos


@pytest.mark.parametrize("replay", [False, True])
def test_adapt_step_streams_match_single_stream(replay):
  """AdaptStepper(two_streams=True) — right-image feature pass on a second stream with deferred BatchNorm running-statistics
  updates, weight gradients on their own stream through the WgradToken nodes — leaves the weights, BN buffers and losses the
  single-stream step leaves (same kernels, same arithmetic; only the order of independent work changes)."""
  from stereonet_b200.adapt import AdaptStepper, make_optimizer
  k, Hh, Ww = 3, 96, 256
  fsd, ssd = O.make_feature_state(k, 11), O.make_stereo_state(22, sharpen=10.0)
  frames = [O.make_stereo_pair(1, Hh, Ww, seed=1000 + i, max_disp_px=40.0) for i in range(3)]
  res = []
  for two in (False, True):
    f = S.FeatureExtractorNetwork(k).to(DEV); s = S.StereoNet(k, 1, 0).to(DEV)
    f.load_state_dict(fsd); s.load_state_dict(ssd)
    st = AdaptStepper(f, s, make_optimizer(f, s, lr=5e-5, fused=True), Hh, Ww, two_streams=two)
    losses = []
    for i in range(2):
      l, r, _ = frames[i]
      rep = tuple(t.to(DEV) for t in frames[2]) if replay else None
      losses.append(st.step(l.to(DEV), r.to(DEV), replay=rep)[0].item())
    torch.cuda.synchronize()
    res.append((losses, {n: v.detach().cpu().clone() for n, v in list(s.state_dict().items()) + list(f.state_dict().items())}))
  (la, wa), (lb, wb) = res
  assert la == lb, (la, lb)
  for n in wa:
    if "num_batches_tracked" in n:
      assert torch.equal(wa[n], wb[n]), n
    elif "running_var" in n or "running_mean" in n:
      # the deferred update rebuilds the batch variance from invstd (1 / invstd^2 - eps): a few ulps of the statistic
      assert (wa[n] - wb[n]).abs().max().item() <= 2e-6 * wa[n].abs().max().item() + 1e-9, n
    else:
      assert torch.equal(wa[n], wb[n]), n


def test_bn_running_update_matches_in_kernel_update():
  """snb_bn_running_update (the deferred half of bn_finalize) against bn_finalize's own in-kernel update of the running buffers."""
  from stereonet_b200 import ops
  torch.manual_seed(0)
  stats = torch.rand(300, 2, 32, device=DEV) * 5.0
  stats[:, 1] += stats[:, 0] ** 2 * 3.0                       # sum z^2 >= (sum z)^2 / n with room to spare
  count = 300 * 128
  a, b = torch.nn.BatchNorm2d(32).to(DEV), torch.nn.BatchNorm2d(32).to(DEV)
  for bn in (a, b):
    with torch.no_grad():
      bn.running_mean.uniform_(-1, 1); bn.running_var.uniform_(0.5, 2.0)
  b.load_state_dict(a.state_dict())
  ops.bn_finalize(stats, count, a)                            # in-kernel update
  _, _, mean, invstd = ops.bn_finalize(stats, count, b, update_running=False)
  ops.bn_running_update(b, mean, invstd, count)
  torch.cuda.synchronize()
  assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 1
  assert (a.running_mean - b.running_mean).abs().max().item() <= 1e-6 * a.running_mean.abs().max().item()
  assert (a.running_var - b.running_var).abs().max().item() <= 2e-6 * a.running_var.abs().max().item()
