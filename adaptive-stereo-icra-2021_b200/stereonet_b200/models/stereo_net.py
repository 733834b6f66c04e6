"""Drop-in replacements for adaptive_stereo/models/stereo_net.py (reference file:line cited per item).

The classes keep the reference's constructor arguments, forward() signatures, output-dict keys, module nesting
(hence state_dict keys/shapes and parameter registration order) — checkpoints load with
strict=True in both directions.  The nn.Conv*/nn.BatchNorm* children are parameter containers only: forward() never
calls them, it hands their tensors to the CUDA kernels of libsnb200.so through stereonet_b200.ops.
"""
import torch
import torch.nn as nn

from .. import ops
from ..autograd import fused


def convbn(in_channels, out_channels, kernel_size, stride, pad, dilation):
  """Parameter container matching convbn() (stereo_net.py:8-18): [0] = Conv2d, [1] = BatchNorm2d."""
  conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                   padding=dilation if dilation > 1 else pad, dilation=dilation)
  return nn.Sequential(conv, nn.BatchNorm2d(out_channels))


def convbn_3d(in_channels, out_channels, kernel_size, stride, pad):
  """Parameter container matching convbn_3d() (stereo_net.py:21-30): [0] = Conv3d, [1] = BatchNorm3d."""
  conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, padding=pad, stride=stride)
  return nn.Sequential(conv, nn.BatchNorm3d(out_channels))


def _to_cl(x, name):
  """Logical NCHW -> physical [B,H,W,C] contiguous (free when the tensor is already channels-last strided)."""
  if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
    raise RuntimeError(f"stereonet_b200: `{name}` must be an fp32 CUDA tensor; this build has no CPU path")
  return x.permute(0, 2, 3, 1).contiguous()


class BasicBlock(nn.Module):
  """stereo_net.py:33-51.  out = x + LeakyReLU(BN(conv1(x))); conv2 exists only for state_dict parity (never run)."""

  def __init__(self, in_channels, out_channels, stride, downsample, pad, dilation):
    super().__init__()
    if in_channels != 32 or out_channels != 32 or stride != 1 or downsample is not None:
      raise NotImplementedError("stereonet_b200.BasicBlock supports the reference's only use: 32->32, stride 1")
    self.conv1 = nn.Sequential(convbn(in_channels, out_channels, 3, stride, pad, dilation),
                               nn.LeakyReLU(negative_slope=0.2, inplace=True))
    self.conv2 = convbn(out_channels, out_channels, 3, 1, pad, dilation)
    self.downsample = downsample
    self.stride = stride
    self.dilation = dilation

  def forward_cl(self, x):
    """x: channels-last [B,H,W,32]."""
    conv, bn = self.conv1[0][0], self.conv1[0][1]
    return fused.conv_bn_lrelu(x, conv, bn, dil=self.dilation, residual=True, training=self.training)

  def forward(self, x):
    return self.forward_cl(_to_cl(x, "x")).permute(0, 3, 1, 2)


class FeatureExtractorNetwork(nn.Module):
  """stereo_net.py:54-85: k x Conv2d(5x5, s2, p2) -> 6 BasicBlocks -> Conv2d(3x3).  rgb [B,3,H,W] -> [B,32,H/2^k,W/2^k]
  (returned as a channels-last-strided NCHW tensor)."""

  def __init__(self, k):
    super().__init__()
    self.k = k
    self.downsample = nn.ModuleList()
    cin = 3
    for _ in range(k):
      self.downsample.append(nn.Conv2d(cin, 32, kernel_size=5, stride=2, padding=2))
      cin = 32
    self.residual_blocks = nn.ModuleList(
        [BasicBlock(32, 32, stride=1, downsample=None, pad=1, dilation=1) for _ in range(6)])
    self.conv_alone = nn.Conv2d(32, 32, kernel_size=3, stride=1, padding=1)

  def forward(self, rgb_img):
    if not (isinstance(rgb_img, torch.Tensor) and rgb_img.is_cuda and rgb_img.dtype == torch.float32):
      raise RuntimeError("stereonet_b200: rgb_img must be an fp32 CUDA tensor; this build has no CPU path")
    if self.k >= 2:
      x = fused.downsample_pair(rgb_img.contiguous(), self.downsample[0], self.downsample[1])
    else:
      x = fused.conv5x5s2_first(rgb_img.contiguous(), self.downsample[0])
    for i in range(2, self.k):
      x = fused.conv_plain(x, self.downsample[i], ksize=5, stride=2)
    for block in self.residual_blocks:
      x = block.forward_cl(x)
    x = fused.conv_plain(x, self.conv_alone, ksize=3, stride=1)
    return x.permute(0, 3, 1, 2)


class EdgeAwareRefinement(nn.Module):
  """stereo_net.py:88-121: upsample coarse disparity, concat with rgb, conv+BN+LReLU, 6 dilated BasicBlocks
  (dilations 1,2,4,8,1,1), Conv2d(32->1), ReLU(upsampled + residual)."""

  def __init__(self, in_channels):
    super().__init__()
    if in_channels != 4:
      raise NotImplementedError("stereonet_b200.EdgeAwareRefinement supports in_channels=4 (disparity + rgb)")
    self.conv2d_feature = nn.Sequential(convbn(in_channels, 32, kernel_size=3, stride=1, pad=1, dilation=1),
                                        nn.LeakyReLU(negative_slope=0.2, inplace=True))
    self.residual_astrous_blocks = nn.ModuleList(
        [BasicBlock(32, 32, stride=1, downsample=None, pad=1, dilation=d) for d in (1, 2, 4, 8, 1, 1)])
    self.conv2d_out = nn.Conv2d(32, 1, kernel_size=3, stride=1, padding=1)

  def forward(self, coarse_disparity, guidance_rgb):
    """coarse_disparity [B,h,w], guidance_rgb [B,3,H,W] -> [B,1,H,W]."""
    return self.forward_with_up(coarse_disparity, guidance_rgb)[0]

  def forward_with_up(self, coarse_disparity, guidance_rgb):
    """-> (refined [B,1,H,W], up [B,H,W]): `up` is the bilinearly upsampled coarse disparity times W / w (stereo_net.py:105-111),
    which the first kernel of the refinement computes anyway."""
    conv, bn = self.conv2d_feature[0][0], self.conv2d_feature[0][1]
    up, x = fused.refine_head(coarse_disparity.contiguous(), guidance_rgb.contiguous(), conv, bn, self.training)
    for block in self.residual_astrous_blocks:
      x = block.forward_cl(x)
    return fused.refine_tail(x, self.conv2d_out, up).unsqueeze(1), up


class DisparityRegression(nn.Module):
  """stereo_net.py:124-134 (imported by evaluation/ood_analysis.py:13): sum_d d * x[:, d].  Inside StereoNet this op is
  fused with conv3d_alone and the softmax (snb_tapsum_softargmin); this class only keeps the import surface."""

  def __init__(self, maxdisp):
    super().__init__()
    self.maxdisp = maxdisp

  def forward(self, x):
    disp = torch.arange(self.maxdisp, dtype=x.dtype, device=x.device).view(1, self.maxdisp, 1, 1)
    return torch.sum(x * disp, 1)


class StereoNet(nn.Module):
  """stereo_net.py:137-207.  forward(left_img, left_features, right_features, side, output_cost_volume=False) -> dict with
  `pred_disp_{side}/{s+k}`, `pred_disp_{side}/{s}` and optionally `cost_volume_{side}/{s+k}`."""

  def __init__(self, k, r, input_scale, maxdisp=192):
    super().__init__()
    self.maxdisp = maxdisp
    self.k = k
    self.r = r
    self.input_scale = input_scale
    self.filter = nn.ModuleList()
    for _ in range(4):
      self.filter.append(nn.Sequential(convbn_3d(32, 32, kernel_size=3, stride=1, pad=1),
                                       nn.LeakyReLU(negative_slope=0.2, inplace=True)))
      # The reference re-creates conv3d_alone inside this loop (stereo_net.py:162); doing the same keeps the RNG
      # stream of a seeded construction identical.
      self.conv3d_alone = nn.Conv3d(32, 1, kernel_size=3, stride=1, padding=1)
    self.edge_aware_refinements = nn.ModuleList([EdgeAwareRefinement(4)])

  def forward(self, left_img, left_features, right_features, side, output_cost_volume=False):
    coarse_max_disp = (self.maxdisp + 1) // pow(2, self.input_scale + self.k)       # stereo_net.py:169
    coarse_scale = self.input_scale + self.k
    if not (isinstance(left_img, torch.Tensor) and left_img.is_cuda and left_img.dtype == torch.float32):
      raise RuntimeError("stereonet_b200: left_img must be an fp32 CUDA tensor; this build has no CPU path")
    left_img = left_img.contiguous()
    H, W = left_img.shape[-2:]
    outputs = {}

    cost = fused.cost_volume(_to_cl(left_features, "left_features"), _to_cl(right_features, "right_features"),
                             coarse_max_disp)                                       # :173-184
    for f in self.filter:                                                           # :185-186
      cost = fused.conv_bn_lrelu(cost, f[0][0], f[0][1], dil=1, residual=False, training=self.training)
    cost_out, pred = fused.conv3d_out_softargmin(cost, self.conv3d_alone, output_cost_volume)   # :187-192

    if output_cost_volume:
      outputs["cost_volume_{}/{}".format(side, coarse_scale)] = cost_out            # :197-198
    if not torch.is_grad_enabled() and W == pred.shape[-1] * (2 ** self.k):
      # the refinement's first kernel upsamples the coarse disparity with the factor W / w (= 2^k here) for its own input:
      # that IS the coarse output of :201-202, so the separate upsampling launch is dropped on the inference path
      refined, up = self.edge_aware_refinements[0].forward_with_up(pred, left_img)
      outputs["pred_disp_{}/{}".format(side, coarse_scale)] = up.unsqueeze(1)
      outputs["pred_disp_{}/{}".format(side, self.input_scale)] = refined
      return outputs
    outputs["pred_disp_{}/{}".format(side, coarse_scale)] = \
        fused.upsample(pred, H, W, float(2 ** self.k)).unsqueeze(1)                # :201-202
    outputs["pred_disp_{}/{}".format(side, self.input_scale)] = \
        self.edge_aware_refinements[0](pred, left_img)                              # :204-205
    return outputs
