"""Loss side of the adaptation step: autograd bindings of the library's fused CUDA kernels.

  monodepth_single_loss   adapt.py:78-86 = LinearWarping.forward(right_to_left=True) (linear_warping.py:18-57) + SSIM / L1 /
                          edge-aware smoothness (loss_functions.py:41-138) + masked mean      -> snb_photo_loss
  khamis_robust_loss      loss_functions.py:6-15 (experience-replay term, adapt.py:343-345)   -> snb_khamis_loss
  feature_contrast_mean   feature_contrast.py:12-23                                           -> snb_feature_contrast, or the map
                          the fused head kernel already wrote next to the cost volume

There is no PyTorch formulation of these in the product: the plain-torch restatement of the reference formulas lives in
oracle/stereonet_oracle.py, where the tests take it from.  CPU tensors raise."""
import torch

from . import ops


def _cuda(t, name):
  if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
    raise RuntimeError(f"stereonet_b200: `{name}` must be an fp32 CUDA tensor; this build has no CPU path")
  return t


class _FusedPhotoLoss(torch.autograd.Function):
  """snb_photo_loss: value and d loss / d disp from one fused pass."""

  @staticmethod
  def forward(ctx, disp, left, right, smooth_w):
    loss, ddisp = ops.photo_loss(left.contiguous(), right.contiguous(), disp.contiguous(), smooth_w)
    ctx.save_for_backward(ddisp)
    return loss[0]

  @staticmethod
  def backward(ctx, gout):
    (ddisp,) = ctx.saved_tensors
    return ddisp * gout, None, None, None


def monodepth_single_loss(left_img, right_img, outputs, scale, smoothness_weight=1e-3):
  """adapt.py:78-86 on the `pred_disp_l/{scale}` entry of a forward's output dict.  Static shapes, no host sync: capturable."""
  pred = _cuda(outputs["pred_disp_l/{}".format(scale)], "pred_disp")
  return _FusedPhotoLoss.apply(pred.squeeze(1), _cuda(left_img, "left_img"), _cuda(right_img, "right_img"), smoothness_weight)


class _FusedKhamisLoss(torch.autograd.Function):
  """snb_khamis_loss: value and d loss / d pred from two launches, no boolean indexing."""

  @staticmethod
  def forward(ctx, pred, gt):
    loss, dpred = ops.khamis_loss(pred.contiguous(), gt.contiguous())
    ctx.save_for_backward(dpred)
    return loss[0]

  @staticmethod
  def backward(ctx, gout):
    (dpred,) = ctx.saved_tensors
    return dpred * gout, None


def khamis_robust_loss(pred_disp, gt_disp):
  """loss_functions.py:6-15: sum_{gt > 0} (sqrt((gt - pred)^2 + 4) / 2 - 1) / max(#valid, 1)."""
  return _FusedKhamisLoss.apply(_cuda(pred_disp, "pred_disp"), _cuda(gt_disp, "gt_disp").reshape(pred_disp.shape))


def feature_contrast_mean(cost_volume):
  """feature_contrast.py:12-23: max_d cost - mean of all but the two largest costs, [B,D,H,W] -> [B,H,W]."""
  _cuda(cost_volume, "cost_volume")
  fcs = getattr(cost_volume, "_snb_fcs", None)            # written by the fused head kernel next to this cost volume
  if fcs is not None and fcs.shape == cost_volume.shape[:1] + cost_volume.shape[2:]:
    return fcs
  with torch.no_grad():
    return ops.feature_contrast(cost_volume.detach().contiguous())
