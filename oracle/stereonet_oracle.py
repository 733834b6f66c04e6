"""ORACLE — test infrastructure, NOT product code.

CPU restatement (plain torch fp32 functional ops + numpy for the exact integer-index
parts) of the StereoNet forward / adaptation hot path of
miloknowles/adaptive-stereo-icra-2021.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product
package (adaptive-stereo-icra-2021_b200/stereonet_b200) never does.

Parity status: the reference's own tests pin NO values for this path
(test/test_stereo_net.py only prints timings), so this oracle is pinned against
outputs of the reference module itself, executed in the build container by
oracle/gen_golden.py and committed under tests/golden/ (see DESIGN.md §3).

Every function cites the reference file:line it restates (paths relative to the
reference checkout).  State is passed as a flat dict keyed exactly like the
reference modules' state_dict().
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2      # stereo_net.py:39,94,159
BN_EPS = 1e-5          # nn.BatchNorm2d/3d default (stereo_net.py:17,29)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _bn(x, sd, prefix, training):
  """nn.BatchNorm2d / nn.BatchNorm3d (stereo_net.py:17, :29) incl. running-stat update."""
  rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
  y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                   training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
  if training and (prefix + ".num_batches_tracked") in sd:
    sd[prefix + ".num_batches_tracked"] += 1
  return y


def _convbn2d(x, sd, prefix, dilation, training):
  """convbn() stereo_net.py:8-18: Conv2d(3x3, stride 1, padding = dilation if dilation > 1 else 1) + BN."""
  pad = dilation if dilation > 1 else 1
  y = F.conv2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], stride=1, padding=pad, dilation=dilation)
  return _bn(y, sd, prefix + ".1", training)


def _basic_block(x, sd, prefix, dilation, training):
  """BasicBlock.forward stereo_net.py:44-51: out = x + LeakyReLU(BN(conv1(x))); conv2 is never called."""
  out = F.leaky_relu(_convbn2d(x, sd, prefix + ".conv1.0", dilation, training), LRELU_SLOPE)
  return x + out


def feature_extractor(sd: Dict[str, torch.Tensor], rgb: torch.Tensor, k: int, training: bool = False):
  """FeatureExtractorNetwork.forward stereo_net.py:79-85 (ctor :54-77)."""
  out = rgb
  for i in range(k):
    out = F.conv2d(out, sd[f"downsample.{i}.weight"], sd[f"downsample.{i}.bias"], stride=2, padding=2)
  for i in range(6):
    out = _basic_block(out, sd, f"residual_blocks.{i}", 1, training)
  return F.conv2d(out, sd["conv_alone.weight"], sd["conv_alone.bias"], stride=1, padding=1)


def cost_volume(left: torch.Tensor, right: torch.Tensor, num_disp: int) -> torch.Tensor:
  """Difference cost volume, stereo_net.py:173-184.

  C[b,c,d,y,x] = L[b,c,y,x] - R[b,c,y,x-d] for x >= d, else 0.  Differentiable restatement
  (pad + slice, no in-place writes) so autograd yields the reference's slice-assign backward.
  """
  B, C, H, W = left.shape
  slabs = []
  for d in range(num_disp):
    if d == 0:
      slabs.append(left - right)
    elif d < W:
      diff = left[..., d:] - right[..., :-d]
      slabs.append(F.pad(diff, (d, 0)))
    else:
      slabs.append(torch.zeros_like(left))
  return torch.stack(slabs, dim=2).contiguous()


def cost_volume_numpy(left: np.ndarray, right: np.ndarray, num_disp: int) -> np.ndarray:
  """Same as cost_volume() in numpy, the literal loop of stereo_net.py:178-182 (bit-exact check)."""
  B, C, H, W = left.shape
  cost = np.zeros((B, C, num_disp, H, W), dtype=np.float32)
  for i in range(num_disp):
    if i > 0:
      cost[:, :, i, :, i:] = left[:, :, :, i:] - right[:, :, :, :-i]
    else:
      cost[:, :, i, :, :] = left - right
  return cost


def cost_filter(sd, cost: torch.Tensor, training: bool = False) -> torch.Tensor:
  """stereo_net.py:185-187: 4x (Conv3d 3^3 p1 + BN3d + LeakyReLU 0.2) then Conv3d(32->1). Returns [B,1,D,H,W]."""
  for i in range(4):
    cost = F.conv3d(cost, sd[f"filter.{i}.0.0.weight"], sd[f"filter.{i}.0.0.bias"], stride=1, padding=1)
    cost = F.leaky_relu(_bn(cost, sd, f"filter.{i}.0.1", training), LRELU_SLOPE)
  return F.conv3d(cost, sd["conv3d_alone.weight"], sd["conv3d_alone.bias"], stride=1, padding=1)


def soft_argmin(cost: torch.Tensor) -> torch.Tensor:
  """stereo_net.py:190-192 + DisparityRegression :124-134. softmax over +cost on dim 1, expectation over d."""
  p = F.softmax(cost, dim=1)
  D = cost.shape[1]
  disp = torch.arange(D, dtype=cost.dtype, device=cost.device).view(1, D, 1, 1)
  return torch.sum(p * disp, 1)


def refinement(sd, coarse: torch.Tensor, rgb: torch.Tensor, training: bool = False, prefix="edge_aware_refinements.0"):
  """EdgeAwareRefinement.forward stereo_net.py:104-121 (dilations [1,2,4,8,1,1], :97)."""
  up = F.interpolate(coarse.unsqueeze(1), size=rgb.shape[-2:], mode="bilinear", align_corners=False)
  up = up * (rgb.shape[-1] / coarse.shape[-1])
  x = torch.cat([up, rgb], dim=1)
  x = F.leaky_relu(_convbn2d(x, sd, prefix + ".conv2d_feature.0", 1, training), LRELU_SLOPE)
  for i, dil in enumerate([1, 2, 4, 8, 1, 1]):
    x = _basic_block(x, sd, f"{prefix}.residual_astrous_blocks.{i}", dil, training)
  res = F.conv2d(x, sd[prefix + ".conv2d_out.weight"], sd[prefix + ".conv2d_out.bias"], stride=1, padding=1)
  return F.relu(up + res)


def stereonet_forward(sd, left_img, left_features, right_features, side: str, k: int, input_scale: int = 0,
                      maxdisp: int = 192, output_cost_volume: bool = False, training: bool = False,
                      extras: Optional[dict] = None):
  """StereoNet.forward stereo_net.py:168-207. `extras` (if a dict) receives intermediates for per-stage parity."""
  coarse_max_disp = (maxdisp + 1) // pow(2, input_scale + k)          # :169
  outputs = {}
  cv = cost_volume(left_features, right_features, coarse_max_disp)    # :173-184
  cost = cost_filter(sd, cv, training).squeeze(1)                     # :185-190
  pred = soft_argmin(cost)                                            # :191-192
  coarse_scale = input_scale + k
  if output_cost_volume:
    outputs[f"cost_volume_{side}/{coarse_scale}"] = cost              # :197-198
  outputs[f"pred_disp_{side}/{coarse_scale}"] = (2 ** k) * F.interpolate(
      pred.unsqueeze(1), size=left_img.shape[-2:], mode="bilinear", align_corners=False)   # :201-202
  outputs[f"pred_disp_{side}/{input_scale}"] = refinement(sd, pred, left_img, training)    # :204-205
  if extras is not None:
    extras["raw_cost_volume"] = cv
    extras["coarse_pred"] = pred
  return outputs


def predict_disparity_left(fsd, ssd, left, right, k, input_scale=0, maxdisp=192, training=False, extras=None):
  """adapt.py:65-75 / train.py:19-22: feature_net(L), feature_net(R), stereo_net(L, fl, fr, 'l', True)."""
  fl = feature_extractor(fsd, left, k, training)
  fr = feature_extractor(fsd, right, k, training)
  if extras is not None:
    extras["left_features"], extras["right_features"] = fl, fr
  return stereonet_forward(ssd, left, fl, fr, "l", k, input_scale, maxdisp, True, training, extras)


# --------------------------------------------------------------------------------------
# loss side of the adaptation step (adjacent to the hot path; defines the backward seed)
# --------------------------------------------------------------------------------------
def linear_warp_right_to_left(img: torch.Tensor, disp: torch.Tensor):
  """LinearWarping.forward(right_to_left=True) linear_warping.py:18-57, incl. its align_corners=False quirk."""
  b, c, h, w = img.shape
  rows, cols = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
  grid = torch.stack([cols, rows], dim=-1).float().to(img.device)
  flow = grid.expand(b, -1, -1, -1).clone()
  flow[..., 0] = flow[..., 0] - disp.permute(0, 2, 3, 1).squeeze(-1)
  flow[..., 0] = (2 * flow[..., 0] / w) - 1.0
  flow[..., 1] = (2 * flow[..., 1] / h) - 1.0
  valid = (flow >= -1.0) * (flow <= 1.0)
  valid = valid[..., 0] * valid[..., 1]
  return F.grid_sample(img, flow, mode="bilinear", padding_mode="border", align_corners=False), valid.unsqueeze(1)


def ssim(x, y):
  """SSIM loss_functions.py:41-72."""
  C1, C2 = 0.01 ** 2, 0.03 ** 2
  mu_x = F.avg_pool2d(x, 3, 1, 1)
  mu_y = F.avg_pool2d(y, 3, 1, 1)
  sigma_x = F.avg_pool2d(x ** 2, 3, 1, 1) - mu_x ** 2
  sigma_y = F.avg_pool2d(y ** 2, 3, 1, 1) - mu_y ** 2
  sigma_xy = F.avg_pool2d(x * y, 3, 1, 1) - mu_x * mu_y
  n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
  d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
  return ((1 - n / d) / 2).clamp(min=0, max=1)


def edge_aware_smoothness(disp, img):
  """monodepth_edge_aware_smoothness_loss loss_functions.py:75-103."""
  gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
  gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
  gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
  giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
  gdx = F.pad(gdx * torch.exp(-gix), (0, 1))
  gdy = F.pad(gdy * torch.exp(-giy), (0, 0, 0, 1))
  return gdx + gdy


def monodepth_single_loss(left_img, right_img, pred_disp, smoothness_weight=1e-3):
  """adapt.py:78-86 + monodepth_loss loss_functions.py:106-138: masked mean of 0.85 SSIM + 0.15 L1 + w * smooth."""
  warped, mask = linear_warp_right_to_left(right_img, pred_disp)
  photo_ssim = ssim(left_img, warped).mean(1, keepdim=True)
  photo_l1 = torch.abs(left_img - warped).mean(1, keepdim=True)
  l_photo = 0.85 * photo_ssim + 0.15 * photo_l1
  mean_disp = pred_disp.mean(2, True).mean(3, True)
  l_smooth = edge_aware_smoothness(pred_disp / (mean_disp + 1e-7), left_img)
  total = l_photo + smoothness_weight * l_smooth
  return total[mask].mean()


def feature_contrast_mean(cost: torch.Tensor) -> torch.Tensor:
  """feature_contrast.py:12-23: sorted_desc[0] - mean(sorted_desc[2:]) over dim 1."""
  s = torch.sort(cost, dim=1, descending=True)[0]
  return s[:, 0] - s[:, 2:].mean(dim=1)


def epe(pred, gt):
  """train.py:103: |pred - gt|[gt > 0].mean()."""
  return torch.abs(pred - gt)[gt > 0].mean()


def khamis_robust_loss(pred_disp: torch.Tensor, gt_disp: torch.Tensor) -> torch.Tensor:
  """khamis_robust_loss loss_functions.py:6-15 (the experience-replay term of adapt.py:343-345):
  sum over gt > 0 of sqrt((gt - pred)^2 + 4) / 2 - 1, divided by max(#valid, 1)."""
  mask = (gt_disp > 0).detach()
  num_valid = max(mask.sum(), 1)
  return torch.sum(torch.sqrt(torch.pow(gt_disp[mask] - pred_disp[mask], 2) + 4) / 2 - 1) / num_valid


D1_THRESHOLDS = (2, 3, 4, 5)          # train.py:106


def eval_metrics(pred_disp: torch.Tensor, gt_disp: torch.Tensor, cost: Optional[torch.Tensor] = None) -> Dict[str, float]:
  """The per-batch body of train.evaluate, train.py:98-110: EPE over gt > 0, D1-all at 2/3/4/5 px, mean FCS."""
  valid = gt_disp > 0
  err = torch.abs(pred_disp - gt_disp)
  m = {"EPE": err[valid].mean().item()}
  for t in D1_THRESHOLDS:
    m[f"D1_all_{t}px"] = ((valid * (err > t)).sum() / float(valid.sum())).item()
  if cost is not None:
    m["FCS"] = feature_contrast_mean(cost).mean().item()
  return m


def validate_ovs(fsd, ssd, pairs, k: int, input_scale: int = 0, maxdisp: int = 192):
  """StateMachine.validate adapt.py:122-142: eval mode, no grad, ONE pair at a time, single-image Monodepth loss.
  `pairs` = iterable of (left [1,3,H,W], right [1,3,H,W]).  Returns the list of per-pair losses."""
  out = []
  with torch.no_grad():
    for left, right in pairs:
      o = predict_disparity_left(fsd, ssd, left, right, k, input_scale, maxdisp, training=False)
      out.append(monodepth_single_loss(left, right, o[f"pred_disp_l/{input_scale}"]).item())
  return out


# --------------------------------------------------------------------------------------
# deterministic weights / inputs shared by golden generation, tests and bench
# --------------------------------------------------------------------------------------
def _u(gen, shape, lo, hi):
  return torch.rand(shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo


def _conv_init(gen, sd, name, cout, cin, *ks):
  fan_in = cin * int(np.prod(ks))
  b = 1.0 / math.sqrt(fan_in)
  sd[name + ".weight"] = _u(gen, (cout, cin, *ks), -b, b) * math.sqrt(3.0)
  sd[name + ".bias"] = _u(gen, (cout,), -b, b)


def _bn_init(gen, sd, name, c=32):
  sd[name + ".weight"] = _u(gen, (c,), 0.6, 1.4)
  sd[name + ".bias"] = _u(gen, (c,), -0.2, 0.2)
  sd[name + ".running_mean"] = _u(gen, (c,), -0.2, 0.2)
  sd[name + ".running_var"] = _u(gen, (c,), 0.5, 1.5)
  sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def make_feature_state(k: int, seed: int) -> Dict[str, torch.Tensor]:
  """Seeded state for FeatureExtractorNetwork(k) with the reference's keys/shapes.

  Not the reference's init distribution: a variance-preserving uniform init with randomised BN
  affine/running stats, so BN folding and batch-stat paths are exercised and signals stay O(1)."""
  g = torch.Generator().manual_seed(seed)
  sd: Dict[str, torch.Tensor] = {}
  cin = 3
  for i in range(k):
    _conv_init(g, sd, f"downsample.{i}", 32, cin, 5, 5)
    cin = 32
  for i in range(6):
    _conv_init(g, sd, f"residual_blocks.{i}.conv1.0.0", 32, 32, 3, 3)
    _bn_init(g, sd, f"residual_blocks.{i}.conv1.0.1")
    _conv_init(g, sd, f"residual_blocks.{i}.conv2.0", 32, 32, 3, 3)
    _bn_init(g, sd, f"residual_blocks.{i}.conv2.1")
  _conv_init(g, sd, "conv_alone", 32, 32, 3, 3)
  return sd


def make_stereo_state(seed: int, sharpen: float = 1.0) -> Dict[str, torch.Tensor]:
  """Seeded state for StereoNet. `sharpen` multiplies conv3d_alone.weight so the filtered cost spans a
  trained-model-like range (random init gives a near-uniform softmax, a vacuous parity test)."""
  g = torch.Generator().manual_seed(seed)
  sd: Dict[str, torch.Tensor] = {}
  for i in range(4):
    _conv_init(g, sd, f"filter.{i}.0.0", 32, 32, 3, 3, 3)
    _bn_init(g, sd, f"filter.{i}.0.1")
  _conv_init(g, sd, "conv3d_alone", 1, 32, 3, 3, 3)
  sd["conv3d_alone.weight"] = sd["conv3d_alone.weight"] * sharpen
  p = "edge_aware_refinements.0"
  _conv_init(g, sd, p + ".conv2d_feature.0.0", 32, 4, 3, 3)
  _bn_init(g, sd, p + ".conv2d_feature.0.1")
  for i in range(6):
    _conv_init(g, sd, f"{p}.residual_astrous_blocks.{i}.conv1.0.0", 32, 32, 3, 3)
    _bn_init(g, sd, f"{p}.residual_astrous_blocks.{i}.conv1.0.1")
    _conv_init(g, sd, f"{p}.residual_astrous_blocks.{i}.conv2.0", 32, 32, 3, 3)
    _bn_init(g, sd, f"{p}.residual_astrous_blocks.{i}.conv2.1")
  _conv_init(g, sd, p + ".conv2d_out", 1, 32, 3, 3)
  # keep the refinement residual small relative to the disparity so ReLU(up + res) is not clamped everywhere
  sd[p + ".conv2d_out.weight"] = sd[p + ".conv2d_out.weight"] * 0.1
  # a trained net keeps a small gain on the (O(100) px) disparity input channel
  sd[p + ".conv2d_feature.0.0.weight"][:, 0] *= 0.02
  return sd


def make_stereo_pair(B: int, H: int, W: int, seed: int, max_disp_px: float = 60.0):
  """Synthetic textured stereo pair with known disparity: left = low-passed noise texture,
  d(y) = 4 + max_disp_px * y/(H-1) px ground-plane ramp, right(x,y) = left(x + d, y) by bilinear sampling
  (restated from LinearWarping(right_to_left=False), linear_warping.py:26-30,47-57). Returns (L, R, gt_disp_l)."""
  g = torch.Generator().manual_seed(seed)
  noise = torch.randn((B, 3, H, W), generator=g)
  low = F.avg_pool2d(F.avg_pool2d(noise, 9, 1, 4), 9, 1, 4) * 6.0
  left = (0.5 + 0.25 * low + 0.1 * torch.randn((B, 3, H, W), generator=g)).clamp(0, 1)
  ramp = 4.0 + max_disp_px * torch.arange(H, dtype=torch.float32) / max(H - 1, 1)
  disp = ramp.view(1, 1, H, 1).expand(B, 1, H, W).contiguous()
  rows, cols = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
  flow = torch.stack([cols, rows], dim=-1).float().expand(B, -1, -1, -1).clone()
  flow[..., 0] = flow[..., 0] + disp.squeeze(1)
  flow[..., 0] = (2 * flow[..., 0] / W) - 1.0
  flow[..., 1] = (2 * flow[..., 1] / H) - 1.0
  right = F.grid_sample(left, flow, mode="bilinear", padding_mode="border", align_corners=False)
  return left.contiguous(), right.contiguous(), disp


def clone_state(sd, requires_grad=False):
  out = {}
  for k, v in sd.items():
    t = v.detach().clone()
    if requires_grad and t.dtype.is_floating_point and "running_" not in k:
      t.requires_grad_(True)
    out[k] = t
  return out


def unused_param(key: str) -> bool:
  """BasicBlock.conv2 is constructed (stereo_net.py:40) but never called -> never receives a gradient."""
  return ".conv2." in key


def adapt_step(fsd, ssd, left, right, k, adam_state, lr=5e-5, input_scale=0, maxdisp=192, clip=True, replay=None,
               er_loss_weight=0.05):
  """One gradient update of the adaptation loop: adapt.py:313-314 (train mode), :328-337 (forward + loss),
  :381-394 (zero_grad, backward, clip_grad_norm_ on stereo_net only, Adam step; params = stereo_net then
  feature_net, :208-210).  Adam restated from torch.optim.Adam defaults (betas .9/.999, eps 1e-8).
  `replay` = (left, right, gt_disp) adds the experience-replay term of the ER / VS+ER modes, adapt.py:339-349,386-388: a SECOND
  full train-mode pass on the replay sample (so BatchNorm running statistics are updated twice) and
  loss += er_loss_weight * khamis_robust_loss(pred_er, gt_er).
  `fsd`/`ssd` hold leaf tensors with requires_grad; updated in place.  Returns (loss, outputs, grads)."""
  outputs = predict_disparity_left(fsd, ssd, left, right, k, input_scale, maxdisp, training=True)
  loss = monodepth_single_loss(left, right, outputs[f"pred_disp_l/{input_scale}"])
  if replay is not None:
    out_er = predict_disparity_left(fsd, ssd, replay[0], replay[1], k, input_scale, maxdisp, training=True)
    loss = loss + er_loss_weight * khamis_robust_loss(out_er[f"pred_disp_l/{input_scale}"], replay[2])
  params = [(("s", n), p) for n, p in ssd.items() if p.requires_grad and not unused_param(n)] + \
           [(("f", n), p) for n, p in fsd.items() if p.requires_grad and not unused_param(n)]
  grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
  gd = {n: (g if g is not None else torch.zeros_like(p)) for (n, p), g in zip(params, grads)}
  if clip:                                                             # adapt.py:391-392
    total = torch.sqrt(sum((g.double() ** 2).sum() for (s, _), g in gd.items() if s == "s")).float()
    coef = torch.clamp(1.0 / (total + 1e-6), max=1.0)
    for n in gd:
      if n[0] == "s":
        gd[n] = gd[n] * coef
  adam_state["step"] = adam_state.get("step", 0) + 1
  t = adam_state["step"]
  b1, b2, eps = 0.9, 0.999, 1e-8
  with torch.no_grad():
    for n, p in params:
      g = gd[n]
      m = adam_state.setdefault(("m", n), torch.zeros_like(p))
      v = adam_state.setdefault(("v", n), torch.zeros_like(p))
      m.mul_(b1).add_(g, alpha=1 - b1)
      v.mul_(b2).addcmul_(g, g, value=1 - b2)
      denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
      p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
  return loss.detach(), outputs, gd
