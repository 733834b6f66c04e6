"""Loss side of the adaptation step, host-side mirror of the reference's L2 code in plain PyTorch ops (it is plain
PyTorch in the reference too and sits OUTSIDE the model drop-in boundary; SURVEY.md §8 f1 lists its fusion as the next
row).  It consumes `pred_disp_*` from the CUDA forward and seeds the CUDA backward through autograd.

  LinearWarping.forward(right_to_left=True)  adaptive_stereo/models/linear_warping.py:18-57
  SSIM / edge-aware smoothness / monodepth   adaptive_stereo/utils/loss_functions.py:41-138
  monodepth_single_loss                      adapt.py:78-86
  khamis_robust_loss                         adaptive_stereo/utils/loss_functions.py:6-15
  feature_contrast_mean                      adaptive_stereo/utils/feature_contrast.py:12-23
"""
import torch
import torch.nn.functional as F


class LinearWarping(torch.nn.Module):
  def __init__(self, height, width, device):
    super().__init__()
    rows, cols = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    self._grid = torch.stack([cols, rows], dim=-1).float().to(device)
    self._height, self._width = height, width

  def forward(self, img, positive_disp, mode="bilinear", right_to_left=True):
    b, c, h, w = img.shape
    assert h == self._height and w == self._width
    flow = self._grid.expand(b, -1, -1, -1).clone()
    shift = positive_disp.permute(0, 2, 3, 1).squeeze(-1)
    flow[..., 0] = flow[..., 0] - shift if right_to_left else flow[..., 0] + shift
    flow[..., 0] = (2 * flow[..., 0] / w) - 1.0
    flow[..., 1] = (2 * flow[..., 1] / h) - 1.0
    valid = (flow >= -1.0) * (flow <= 1.0)
    valid = valid[..., 0] * valid[..., 1]
    # the reference relies on grid_sample's default align_corners=False (a -0.5 px quirk)
    return F.grid_sample(img, flow, mode=mode, padding_mode="border", align_corners=False), valid.unsqueeze(1)


def ssim(x, y):
  c1, c2 = 0.01 ** 2, 0.03 ** 2
  mu_x, mu_y = F.avg_pool2d(x, 3, 1, 1), F.avg_pool2d(y, 3, 1, 1)
  sigma_x = F.avg_pool2d(x ** 2, 3, 1, 1) - mu_x ** 2
  sigma_y = F.avg_pool2d(y ** 2, 3, 1, 1) - mu_y ** 2
  sigma_xy = F.avg_pool2d(x * y, 3, 1, 1) - mu_x * mu_y
  n = (2 * mu_x * mu_y + c1) * (2 * sigma_xy + c2)
  d = (mu_x ** 2 + mu_y ** 2 + c1) * (sigma_x + sigma_y + c2)
  return ((1 - n / d) / 2).clamp(min=0, max=1)


def edge_aware_smoothness(disp, img):
  gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
  gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
  gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
  giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
  return F.pad(gdx * torch.exp(-gix), (0, 1)) + F.pad(gdy * torch.exp(-giy), (0, 0, 0, 1))


def monodepth_loss(pred_disp, true_img, warped_img, smoothness_weight=0.001):
  photo_ssim = ssim(true_img, warped_img).mean(1, keepdim=True)
  photo_l1 = torch.abs(true_img - warped_img).mean(1, keepdim=True)
  l_photo = 0.85 * photo_ssim + 0.15 * photo_l1
  mean_disp = pred_disp.mean(2, True).mean(3, True)
  l_smooth = edge_aware_smoothness(pred_disp / (mean_disp + 1e-7), true_img)
  return l_photo + smoothness_weight * l_smooth, photo_l1, photo_ssim, l_smooth


def monodepth_single_loss(left_img, right_img, outputs, warper, scale, static_shapes=False):
  """adapt.py:78-86.  static_shapes=True computes the same masked mean as sum(loss*mask)/sum(mask) instead of boolean
  indexing, so the step has no data-dependent shapes / host syncs and can be captured in a CUDA graph."""
  key = "pred_disp_l/{}".format(scale)
  left_warped, mask = warper(right_img, outputs[key], right_to_left=True)
  loss = monodepth_loss(outputs[key], left_img, left_warped, smoothness_weight=1e-3)[0]
  if static_shapes:
    m = mask.to(loss.dtype)
    return (loss * m).sum() / m.sum()
  return loss[mask].mean()


class _FusedPhotoLoss(torch.autograd.Function):
  """snb_photo_loss: value and d loss / d disp from one fused pass."""

  @staticmethod
  def forward(ctx, disp, left, right, smooth_w):
    from . import ops
    loss, ddisp = ops.photo_loss(left.contiguous(), right.contiguous(), disp.contiguous(), smooth_w)
    ctx.save_for_backward(ddisp)
    return loss[0]

  @staticmethod
  def backward(ctx, gout):
    (ddisp,) = ctx.saved_tensors
    return ddisp * gout, None, None, None


def monodepth_single_loss_fused(left_img, right_img, outputs, scale):
  """Same value and gradient as monodepth_single_loss, computed by the library's fused CUDA kernels."""
  pred = outputs["pred_disp_l/{}".format(scale)]
  return _FusedPhotoLoss.apply(pred.squeeze(1), left_img, right_img, 1e-3)


def khamis_robust_loss(pred_disp, gt_disp):
  mask = (gt_disp > 0).detach()
  num_valid = max(mask.sum(), 1)
  return torch.sum(torch.sqrt(torch.pow(gt_disp[mask] - pred_disp[mask], 2) + 4) / 2 - 1) / num_valid


class _FusedKhamisLoss(torch.autograd.Function):
  """snb_khamis_loss: value and d loss / d pred from two launches, no boolean indexing."""

  @staticmethod
  def forward(ctx, pred, gt):
    from . import ops
    loss, dpred = ops.khamis_loss(pred.contiguous(), gt.contiguous())
    ctx.save_for_backward(dpred)
    return loss[0]

  @staticmethod
  def backward(ctx, gout):
    (dpred,) = ctx.saved_tensors
    return dpred * gout, None


def khamis_robust_loss_fused(pred_disp, gt_disp):
  """Same value and gradient as khamis_robust_loss, static shapes (CUDA-graph capturable)."""
  return _FusedKhamisLoss.apply(pred_disp, gt_disp)


def feature_contrast_mean(cost_volume):
  """feature_contrast.py:12-23.  CUDA tensors use the library's single-pass kernel (no sort); the torch expression below is
  the reference formulation (kept for CPU tensors in tests)."""
  if cost_volume.is_cuda:
    from . import ops
    fcs = getattr(cost_volume, "_snb_fcs", None)          # written by the fused head kernel next to this cost volume
    if fcs is not None and fcs.shape == cost_volume.shape[:1] + cost_volume.shape[2:]:
      return fcs
    with torch.no_grad():
      return ops.feature_contrast(cost_volume.detach().contiguous())
  with torch.no_grad():
    s = torch.sort(cost_volume, dim=1, descending=True)[0]
    return s[:, 0] - s[:, 2:].mean(dim=1)
