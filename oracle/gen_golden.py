"""Generate golden vectors by executing the UNMODIFIED reference module (build container only).

Usage (in the build container, where /root/reference exists):
    python oracle/gen_golden.py [base] [big] [rows] [layout]     # writes tests/golden/* (default: all groups)

The reference's `forward` hard-codes `.cuda()` (adaptive_stereo/models/stereo_net.py:129,177), so on this
GPU-less host `Tensor.cuda` / `Module.cuda` are shimmed to identity.  The reference modules are
loaded by file path; weights and inputs come from the seeded generators in oracle/stereonet_oracle.py so the
fixtures only need to hold OUTPUTS (plus checksums of the regenerated inputs to detect RNG drift).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import stereonet_oracle as O  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name: (B, H, W, k, input_scale, sharpen, with_train)
TRAIN_SHARPEN_RATIO = 0.25

CASES = {
  "k3_small_dgtw":   dict(B=1, H=64, W=128, k=3, s=0, sharpen=1.0, train=False),    # D'=24 > W'=16
  "k3_b2_sharp":     dict(B=2, H=96, W=256, k=3, s=0, sharpen=40.0, train=True),
  "k4_sharp":        dict(B=1, H=64, W=256, k=4, s=0, sharpen=40.0, train=False),   # D'=12
  "k3_ragged":       dict(B=1, H=68, W=120, k=3, s=0, sharpen=40.0, train=True),    # 68->34->17->9 rows
  "k3_scale1":       dict(B=1, H=48, W=96,  k=3, s=1, sharpen=40.0, train=False),   # D'=12 via input_scale
}


def _load(name, rel):
  spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def load_reference():
  torch.Tensor.cuda = lambda self, *a, **k: self
  torch.nn.Module.cuda = lambda self, *a, **k: self
  sn = _load("ref_stereo_net", "adaptive_stereo/models/stereo_net.py")
  lw = _load("ref_linear_warping", "adaptive_stereo/models/linear_warping.py")
  # loss_functions imports nothing from the package at module level
  lf = _load("ref_loss_functions", "adaptive_stereo/utils/loss_functions.py")
  fc = _load("ref_feature_contrast", "adaptive_stereo/utils/feature_contrast.py")
  return sn, lw, lf, fc


def summarize(t: torch.Tensor):
  t = t.detach().double().flatten()
  return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().sqrt().item()], dtype=np.float64)


def run_case(name, cfg, sn, lw, lf, fc):
  B, H, W, k, s = cfg["B"], cfg["H"], cfg["W"], cfg["k"], cfg["s"]
  fsd = O.make_feature_state(k, seed=11)
  ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"])
  left, right, gt = O.make_stereo_pair(B, H, W, seed=1000, max_disp_px=min(60.0, W / 4))
  out = {"in_left_sum": summarize(left), "in_right_sum": summarize(right),
         "w_feat_sum": summarize(torch.cat([v.flatten().float() for v in fsd.values()])),
         "w_stereo_sum": summarize(torch.cat([v.flatten().float() for v in ssd.values()]))}

  fnet = sn.FeatureExtractorNetwork(k)
  snet = sn.StereoNet(k, 1, s, maxdisp=192)
  fnet.load_state_dict(fsd, strict=True)
  snet.load_state_dict(ssd, strict=True)

  # ---- eval / no-grad forward (train.process_batch, train.py:19-22)
  fnet.eval(); snet.eval()
  with torch.no_grad():
    fl, fr = fnet(left), fnet(right)
    o = snet(left, fl, fr, "l", output_cost_volume=True)
  out["eval/left_features"] = fl.numpy()
  out["eval/right_features"] = fr.numpy()
  for key, v in o.items():
    out["eval/" + key] = v.numpy()
  out["eval/fcs"] = fc.feature_contrast_mean(o[f"cost_volume_l/{s + k}"]).mean().numpy()
  out["eval/epe"] = torch.abs(o[f"pred_disp_l/{s}"] - gt)[gt > 0].mean().numpy()

  if cfg["train"]:
    # train-mode BN renormalises the filter activations, so the same conv3d_alone gain gives a ~3x wider cost range
    # than in eval mode; use a milder gain to stay in a trained-model-like range (about [-22, 12]).
    ssd_t = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"] * TRAIN_SHARPEN_RATIO)
    snet.load_state_dict(ssd_t, strict=True)
    # ---- one adaptation step exactly as adapt.py:313-337,381-394 (Adam lr 5e-5, clip on stereo_net only)
    fnet.train(); snet.train()
    opt = torch.optim.Adam([{"params": snet.parameters()}, {"params": fnet.parameters()}], lr=5e-5)
    warper = lw.LinearWarping(H, W, torch.device("cpu"))
    fl, fr = fnet(left), fnet(right)
    o = snet(left, fl, fr, "l", output_cost_volume=True)
    lw_img, mask = warper(right, o[f"pred_disp_l/{s}"], right_to_left=True)
    loss = lf.monodepth_loss(o[f"pred_disp_l/{s}"], left, lw_img, smoothness_weight=1e-3)[0][mask].mean()
    opt.zero_grad()
    loss.backward()
    out["train/loss"] = loss.detach().numpy()
    out["train/left_features"] = fl.detach().numpy()
    for key, v in o.items():
      out["train/" + key] = v.detach().numpy()
    for tag, net in (("s", snet), ("f", fnet)):
      for n, p in net.named_parameters():
        if p.grad is None:
          out[f"grad_none/{tag}/{n}"] = np.array(1)
        else:
          out[f"grad_sum/{tag}/{n}"] = summarize(p.grad)
          if p.grad.numel() <= 1024:
            out[f"grad/{tag}/{n}"] = p.grad.numpy().copy()
    # a few full-size gradients for direct comparison
    out["grad/s/filter.0.0.0.weight"] = snet.filter[0][0][0].weight.grad.numpy().copy()
    out["grad/f/downsample.0.weight"] = fnet.downsample[0].weight.grad.numpy().copy()
    out["grad/f/conv_alone.weight"] = fnet.conv_alone.weight.grad.numpy().copy()
    out["train/grad_norm_stereo"] = torch.nn.utils.clip_grad_norm_(snet.parameters(), 1.0).numpy()
    opt.step()
    for tag, net in (("s", snet), ("f", fnet)):
      for n, v in net.state_dict().items():
        if "running_" in n or "num_batches" in n:
          out[f"post/{tag}/{n}"] = v.numpy().copy()
        else:
          out[f"post_sum/{tag}/{n}"] = summarize(v)
    out["post/s/conv3d_alone.weight"] = snet.conv3d_alone.weight.detach().numpy().copy()
    out["post/f/conv_alone.bias"] = fnet.conv_alone.bias.detach().numpy().copy()

  os.makedirs(OUT, exist_ok=True)
  np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
  c = out[f"eval/cost_volume_l/{s + k}"]
  print(f"{name}: cost range [{c.min():.2f}, {c.max():.2f}] fcs {out['eval/fcs']:.3f} epe {out['eval/epe']:.3f} "
        f"disp mean {out[f'eval/pred_disp_l/{s}'].mean():.3f} std {out[f'eval/pred_disp_l/{s}'].std():.3f}"
        + (f" loss {out['train/loss']:.5f} gnorm {out['train/grad_norm_stereo']:.4f} train cost range "
           f"[{out[f'train/cost_volume_l/{s + k}'].min():.1f}, {out[f'train/cost_volume_l/{s + k}'].max():.1f}]" if cfg["train"] else ""))


# ---- round 2: the reference's own test / timing configurations at full size (test/test_stereo_net.py:17-22,46-72,
# evaluation/stereonet_timing.py:22-41 and every experiments/adaptation/*.sh: k = 4), eval mode.  Full-resolution tensors are
# stored subsampled (every SUB-th pixel of the upsampled coarse map, the whole refined map) to keep the fixtures small.
BIG_CASES = {
  "T_320x960_k3":   dict(B=1, H=320, W=960, k=3, s=0, sharpen=40.0),     # BASELINE.json configs[0]
  "k4_320x960":     dict(B=1, H=320, W=960, k=4, s=0, sharpen=40.0),     # the experiments' k = 4
  "k3s1_160x480":   dict(B=1, H=160, W=480, k=3, s=1, sharpen=40.0),     # input_scale = 1 (half-resolution input)
}
SUB = 4


def run_big_case(name, cfg, sn, lw, lf, fc):
  B, H, W, k, s = cfg["B"], cfg["H"], cfg["W"], cfg["k"], cfg["s"]
  fsd = O.make_feature_state(k, seed=11)
  ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"])
  left, right, gt = O.make_stereo_pair(B, H, W, seed=1000, max_disp_px=min(60.0, W / 4))
  out = {"in_left_sum": summarize(left), "in_right_sum": summarize(right)}
  fnet = sn.FeatureExtractorNetwork(k); snet = sn.StereoNet(k, 1, s, maxdisp=192)
  fnet.load_state_dict(fsd, strict=True); snet.load_state_dict(ssd, strict=True)
  fnet.eval(); snet.eval()
  with torch.no_grad():
    fl, fr = fnet(left), fnet(right)
    o = snet(left, fl, fr, "l", output_cost_volume=True)
  out["eval/left_features_sum"] = summarize(fl)
  out["eval/left_features_sub"] = fl[..., ::SUB, ::SUB].numpy().copy()
  out[f"eval/cost_volume_l/{s + k}"] = o[f"cost_volume_l/{s + k}"].numpy()
  out[f"eval/pred_disp_l/{s + k}_sub"] = o[f"pred_disp_l/{s + k}"][..., ::SUB, ::SUB].numpy().copy()
  out[f"eval/pred_disp_l/{s}"] = o[f"pred_disp_l/{s}"].numpy()
  out["eval/fcs"] = fc.feature_contrast_mean(o[f"cost_volume_l/{s + k}"]).mean().numpy()
  pred = o[f"pred_disp_l/{s}"]
  valid = gt > 0
  out["eval/epe"] = torch.abs(pred - gt)[valid].mean().numpy()                                      # train.py:103
  out["eval/d1_all"] = np.array([((valid * (torch.abs(pred - gt) > ot)).sum() / float(valid.sum())).item()
                                 for ot in [2, 3, 4, 5]])                                           # train.py:106-107
  np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
  c = out[f"eval/cost_volume_l/{s + k}"]
  print(f"{name}: cost range [{c.min():.2f}, {c.max():.2f}] fcs {out['eval/fcs']:.3f} epe {out['eval/epe']:.3f} d1 {out['eval/d1_all']}")


def run_rows_case(sn, lw, lf, fc):
  """SURVEY.md section 8 rows f3 / f4, from the reference's own functions: khamis_robust_loss (loss_functions.py:6-15) with its
  gradient, the evaluate() metrics (train.py:98-110, restated line by line: train.py cannot be imported here), the
  one-pair-at-a-time OVS validation loop (adapt.py:122-142) and one two-pass experience-replay Adam step (adapt.py:328-349,381-394)."""
  out = {}
  H, W, k, s = 96, 256, 3, 0
  fsd = O.make_feature_state(k, seed=11)
  ssd = O.make_stereo_state(seed=22, sharpen=10.0)
  fnet = sn.FeatureExtractorNetwork(k); snet = sn.StereoNet(k, 1, s, maxdisp=192)
  fnet.load_state_dict(fsd, strict=True); snet.load_state_dict(ssd, strict=True)
  warper = lw.LinearWarping(H, W, torch.device("cpu"))

  # ---- f3: Khamis loss on synthetic (pred, gt) incl. invalid pixels, value and gradient
  g = torch.Generator().manual_seed(3)
  gt = torch.rand((2, 1, 64, 130), generator=g) * 60 + 0.5
  gt[torch.rand((2, 1, 64, 130), generator=g) >= 0.6] = 0.0
  pred = (gt + 4.0 * torch.randn((2, 1, 64, 130), generator=g)).abs().requires_grad_()
  loss = lf.khamis_robust_loss(pred, gt)
  loss.backward()
  out["khamis/gt"] = gt.numpy(); out["khamis/pred"] = pred.detach().numpy()
  out["khamis/loss"] = loss.detach().numpy(); out["khamis/dpred"] = pred.grad.numpy().copy()
  out["khamis/loss_all_invalid"] = lf.khamis_robust_loss(pred.detach(), torch.zeros_like(gt)).numpy()

  # ---- f4: evaluate() over 3 batches of 2 pairs (train.py:74-126), eval mode
  fnet.eval(); snet.eval()
  EPEs, D1s, FCSs = torch.zeros(3), torch.zeros(3, 4), torch.zeros(3)
  with torch.no_grad():
    for i in range(3):
      l, r, gtd = O.make_stereo_pair(2, H, W, seed=3000 + i, max_disp_px=40.0)
      o = snet(l, fnet(l), fnet(r), "l", output_cost_volume=True)
      pd = o[f"pred_disp_l/{s}"]
      valid = gtd > 0
      EPEs[i] = torch.abs(pd - gtd)[valid].mean()
      for oi, ot in enumerate([2, 3, 4, 5]):
        D1s[i, oi] = (valid * (torch.abs(pd - gtd) > ot)).sum() / float(valid.sum())
      FCSs[i] = fc.feature_contrast_mean(o[f"cost_volume_l/{s + k}"]).mean()
  out["evaluate/EPE"] = EPEs.mean().numpy(); out["evaluate/D1_all"] = D1s.mean(dim=0).numpy(); out["evaluate/FCS"] = FCSs.mean().numpy()
  out["evaluate/EPE_per_batch"] = EPEs.numpy(); out["evaluate/D1_per_batch"] = D1s.numpy()

  # ---- f4: StateMachine.validate loop over 5 OVS pairs (adapt.py:131-139)
  vals = []
  with torch.no_grad():
    for i in range(5):
      l, r, _ = O.make_stereo_pair(1, H, W, seed=2000 + i, max_disp_px=40.0)
      o = snet(l, fnet(l), fnet(r), "l", output_cost_volume=True)
      lw_img, mask = warper(r, o[f"pred_disp_l/{s}"], right_to_left=True)
      vals.append(lf.monodepth_loss(o[f"pred_disp_l/{s}"], l, lw_img, smoothness_weight=1e-3)[0][mask].mean().item())
  out["validate/losses"] = np.array(vals)

  # ---- f3: one ER step (two train-mode passes, adapt.py:328-349,381-394), Adam lr 5e-5, er_loss_weight 0.05
  fnet.train(); snet.train()
  opt = torch.optim.Adam([{"params": snet.parameters()}, {"params": fnet.parameters()}], lr=5e-5)
  l, r, _ = O.make_stereo_pair(1, H, W, seed=1000, max_disp_px=40.0)
  rl, rr, rgt = O.make_stereo_pair(1, H, W, seed=1001, max_disp_px=40.0)
  o = snet(l, fnet(l), fnet(r), "l", output_cost_volume=True)
  lw_img, mask = warper(r, o[f"pred_disp_l/{s}"], right_to_left=True)
  l_mono = lf.monodepth_loss(o[f"pred_disp_l/{s}"], l, lw_img, smoothness_weight=1e-3)[0][mask].mean()
  o_er = snet(rl, fnet(rl), fnet(rr), "l", output_cost_volume=True)
  l_er = lf.khamis_robust_loss(o_er[f"pred_disp_l/{s}"], rgt)
  opt.zero_grad()
  backprop = l_mono + 0.05 * l_er
  backprop.backward()
  out["er/loss_mono"] = l_mono.detach().numpy(); out["er/loss_replay"] = l_er.detach().numpy(); out["er/loss"] = backprop.detach().numpy()
  out["er/pred_replay"] = o_er[f"pred_disp_l/{s}"].detach().numpy()
  for tag, net in (("s", snet), ("f", fnet)):
    for n, p in net.named_parameters():
      if p.grad is not None:
        out[f"er/grad_sum/{tag}/{n}"] = summarize(p.grad)
  out["er/grad/s/conv3d_alone.weight"] = snet.conv3d_alone.weight.grad.numpy().copy()
  out["er/grad/f/conv_alone.bias"] = fnet.conv_alone.bias.grad.numpy().copy()
  out["er/grad_norm_stereo"] = torch.nn.utils.clip_grad_norm_(snet.parameters(), 1.0).numpy()
  opt.step()
  for tag, net in (("s", snet), ("f", fnet)):
    for n, v in net.state_dict().items():
      if "running_" in n or "num_batches" in n:
        out[f"er/post/{tag}/{n}"] = v.numpy().copy()
      else:
        out[f"er/post_sum/{tag}/{n}"] = summarize(v)
  out["er/post/s/conv3d_alone.weight"] = snet.conv3d_alone.weight.detach().numpy().copy()
  np.savez_compressed(os.path.join(OUT, "rows_f3_f4.npz"), **out)
  print(f"rows_f3_f4: khamis {out['khamis/loss']:.5f} evaluate EPE {out['evaluate/EPE']:.4f} D1 {out['evaluate/D1_all']} "
        f"validate {out['validate/losses']} er loss {out['er/loss']:.5f} (mono {out['er/loss_mono']:.5f} replay {out['er/loss_replay']:.4f}) "
        f"gnorm {out['er/grad_norm_stereo']:.4f}")


def run_layout_case(sn):
  """state_dict layout and seeded-init fingerprint of the reference classes (adapt.py:29 / train.py:141 seed 123, construction
  order feature_net then stereo_net, train.py:162-163): names, shapes, dtypes and (sum, |sum|, l2) per tensor, k = 3 and k = 4.
  tests/test_state_dict.py holds the drop-in classes to it on every box (the reference checkout only exists here)."""
  import json
  out = {}
  for k in (3, 4):
    torch.manual_seed(123)
    f, s = sn.FeatureExtractorNetwork(k), sn.StereoNet(k, 1, 0, maxdisp=192)
    for tag, net in (("feature_net", f), ("stereo_net", s)):
      out[f"k{k}/{tag}"] = {
        "state_dict": [[n, list(v.shape), str(v.dtype), summarize(v.float()).tolist()] for n, v in net.state_dict().items()],
        "parameters": [n for n, _ in net.named_parameters()],
      }
  with open(os.path.join(OUT, "state_layout.json"), "w") as fh:
    json.dump(out, fh)
  print("state_layout:", {k_: len(v["state_dict"]) for k_, v in out.items()})


if __name__ == "__main__":
  torch.manual_seed(123)
  torch.set_num_threads(8)
  mods = load_reference()
  which = sys.argv[1:] or ["base", "big", "rows", "layout"]
  os.makedirs(OUT, exist_ok=True)
  if "base" in which:
    for name, cfg in CASES.items():
      run_case(name, cfg, *mods)
  if "big" in which:
    for name, cfg in BIG_CASES.items():
      run_big_case(name, cfg, *mods)
  if "rows" in which:
    run_rows_case(*mods)
  if "layout" in which:
    run_layout_case(mods[0])
