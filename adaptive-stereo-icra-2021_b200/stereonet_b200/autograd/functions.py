"""torch.autograd.Function wrappers for the adaptation step (adapt.py:381-394).  Forward and backward both run in the
library's CUDA kernels; torch only carries the graph, the saved tensors and a few layout views."""
import torch

from .. import ops
from . import fused


def _c(t):
  return t if t.is_contiguous() else t.contiguous()


class ConvBnLrelu(torch.autograd.Function):
  """[x +] LeakyReLU(BN(conv3x3(x) + b)), train (batch stats) or eval (running stats)."""

  @staticmethod
  def forward(ctx, x, w, b, gamma, beta, conv, bn, dil, residual, training):
    g = ops.geom(x.shape, 3, stride=1, dil=dil)
    z, stats = fused.conv3x3_c32(x, conv, g, bias=b.detach(), want_stats=training)
    if training:
      scale, shift, mean, invstd = fused.bn_finalize(stats, z.numel() // 32, bn)
    else:
      scale, shift = fused.bn_fold(bn)
      mean, invstd = bn.running_mean, fused.bn_invstd(bn)
    y = ops.bn_apply(z, scale, shift, residual=x if residual else None, lrelu=True)
    ctx.save_for_backward(x, z, scale, shift, mean, invstd)
    ctx.conv, ctx.dil, ctx.residual, ctx.training, ctx.wshape = conv, dil, residual, training, tuple(w.shape)
    return y

  @staticmethod
  def backward(ctx, dy):
    x, z, scale, shift, mean, invstd = ctx.saved_tensors
    dy = _c(dy)
    dz, dgamma, dbeta, dbias = ops.bn_lrelu_bwd(z, dy, scale, shift, mean, invstd, ctx.training, lrelu=True)
    g = ops.geom(x.shape, 3, stride=1, dil=ctx.dil)
    dx = None
    if ctx.needs_input_grad[0]:
      dx, _ = fused.conv3x3_c32_dgrad(dz, ctx.conv, g, residual=dy if ctx.residual else None)
    dw = fused.conv3x3_c32_wgrad(x, dz, g, ctx.wshape)
    return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None


def _on_this_stream(*tensors):
  """Tensors allocated on another stream are about to be read by kernels of the current one: tell the caching allocator, or
  their memory could be handed out again on the owning stream while these kernels still run."""
  cur = torch.cuda.current_stream(tensors[0].device)
  for t in tensors:
    t.record_stream(cur)


class WgradToken(torch.autograd.Function):
  """Carries the weight of a conv into the graph on its OWN stream.  forward returns a zero-stride token shaped like the conv's
  output (one float of storage, no kernel); ConvBnLreluSplit returns dz as the token's gradient, so THIS node's backward — which
  autograd runs on the stream this forward ran on, with the proper event waits — computes the weight gradient next to the data
  path instead of inside it: the tensor-bound wgrad kernel overlaps the HBM-bound BatchNorm backward of the next layer."""

  @staticmethod
  def forward(ctx, w, x, dil):
    ctx.save_for_backward(x)
    ctx.dil, ctx.wshape = dil, tuple(w.shape)
    return fused.wgrad_token(x.shape, x.device)

  @staticmethod
  def backward(ctx, dz):
    (x,) = ctx.saved_tensors
    dz = _c(dz)
    _on_this_stream(x, dz)
    g = ops.geom(x.shape, 3, stride=1, dil=ctx.dil)
    return fused.conv3x3_c32_wgrad(x, dz, g, ctx.wshape), None, None


class ConvBnLreluSplit(torch.autograd.Function):
  """ConvBnLrelu with the weight gradient split off into WgradToken (same kernels, same arithmetic)."""

  @staticmethod
  def forward(ctx, x, tok, b, gamma, beta, conv, bn, dil, residual, training):
    g = ops.geom(x.shape, 3, stride=1, dil=dil)
    z, stats = fused.conv3x3_c32(x, conv, g, bias=b.detach(), want_stats=training)
    if training:
      scale, shift, mean, invstd = fused.bn_finalize(stats, z.numel() // 32, bn)
    else:
      scale, shift = fused.bn_fold(bn)
      mean, invstd = bn.running_mean, fused.bn_invstd(bn)
    y = ops.bn_apply(z, scale, shift, residual=x if residual else None, lrelu=True)
    ctx.save_for_backward(x, z, scale, shift, mean, invstd)
    ctx.conv, ctx.dil, ctx.residual, ctx.training = conv, dil, residual, training
    return y

  @staticmethod
  def backward(ctx, dy):
    x, z, scale, shift, mean, invstd = ctx.saved_tensors
    dy = _c(dy)
    dz, dgamma, dbeta, dbias = ops.bn_lrelu_bwd(z, dy, scale, shift, mean, invstd, ctx.training, lrelu=True)
    g = ops.geom(x.shape, 3, stride=1, dil=ctx.dil)
    dx = None
    if ctx.needs_input_grad[0]:
      dx, _ = fused.conv3x3_c32_dgrad(dz, ctx.conv, g, residual=dy if ctx.residual else None)
    return dx, dz, dbias, dgamma, dbeta, None, None, None, None, None


def conv_bn_lrelu_autograd(x, conv, bn, dil, residual, training):
  ws = fused.WGRAD_STREAM
  if ws is None or not conv.weight.requires_grad:
    return ConvBnLrelu.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, conv, bn, dil, residual, training)
  with torch.cuda.stream(ws):
    tok = WgradToken.apply(conv.weight, x.detach(), dil)
  return ConvBnLreluSplit.apply(x, tok, conv.bias, bn.weight, bn.bias, conv, bn, dil, residual, training)


class ConvC32(torch.autograd.Function):
  """Conv2d 32->32 with bias only: 5x5 stride 2 (downsample[1:]) or 3x3 stride 1 (conv_alone)."""

  @staticmethod
  def forward(ctx, x, w, b, conv, ksize, stride):
    g = ops.geom(x.shape, ksize, stride=stride, dil=1, pad=ksize // 2)
    if ksize == 3 and stride == 1:
      y, _ = fused.conv3x3_c32(x, conv, g, bias=b.detach())
    elif ksize == 5 and stride == 2:
      y = fused.conv5x5s2_c32(x, conv, b.detach())
    else:
      y, _ = ops.conv_c32(x, fused.wprep(conv), g, bias=b.detach())
    ctx.save_for_backward(x)
    ctx.conv, ctx.ksize, ctx.stride, ctx.wshape = conv, ksize, stride, tuple(w.shape)
    return y

  @staticmethod
  def backward(ctx, dy):
    (x,) = ctx.saved_tensors
    dy = _c(dy)
    g = ops.geom(x.shape, ctx.ksize, stride=ctx.stride, dil=1, pad=ctx.ksize // 2)
    dx = None
    if ctx.needs_input_grad[0]:
      if ctx.ksize == 3 and ctx.stride == 1:
        dx, _ = fused.conv3x3_c32_dgrad(dy, ctx.conv, g)
      elif ctx.ksize == 5 and ctx.stride == 2:
        dx = fused.conv5x5s2_c32_dgrad(dy, ctx.conv, x.shape[1], x.shape[2])
      else:
        gt = ops.geom_transposed(g)
        dx, _ = ops.conv_c32(dy, fused.wprep(ctx.conv, 2), gt)
    if ctx.ksize == 3 and ctx.stride == 1:
      dw = fused.conv3x3_c32_wgrad(x, dy, g, ctx.wshape)
    else:
      dw = ops.conv_c32_wgrad(x, dy, g, ctx.wshape)
    return dx, dw, ops.channel_sum(dy), None, None, None


class WgradTokenC32(torch.autograd.Function):
  """WgradToken for ConvC32 (bias-only 32->32 convs: 5x5 stride 2, 3x3): see WgradToken."""

  @staticmethod
  def forward(ctx, w, x, ksize, stride):
    g = ops.geom(x.shape, ksize, stride=stride, dil=1, pad=ksize // 2)
    ctx.save_for_backward(x)
    ctx.ksize, ctx.stride, ctx.wshape = ksize, stride, tuple(w.shape)
    oshape = tuple(ops.out_shape(g, False))
    return fused.wgrad_token(oshape, x.device)

  @staticmethod
  def backward(ctx, dy):
    (x,) = ctx.saved_tensors
    dy = _c(dy)
    _on_this_stream(x, dy)
    g = ops.geom(x.shape, ctx.ksize, stride=ctx.stride, dil=1, pad=ctx.ksize // 2)
    if ctx.ksize == 3 and ctx.stride == 1:
      return fused.conv3x3_c32_wgrad(x, dy, g, ctx.wshape), None, None, None
    return ops.conv_c32_wgrad(x, dy, g, ctx.wshape), None, None, None


class ConvC32Split(torch.autograd.Function):
  """ConvC32 with the weight gradient split off into WgradTokenC32."""

  @staticmethod
  def forward(ctx, x, tok, b, conv, ksize, stride):
    g = ops.geom(x.shape, ksize, stride=stride, dil=1, pad=ksize // 2)
    if ksize == 3 and stride == 1:
      y, _ = fused.conv3x3_c32(x, conv, g, bias=b.detach())
    elif ksize == 5 and stride == 2:
      y = fused.conv5x5s2_c32(x, conv, b.detach())
    else:
      y, _ = ops.conv_c32(x, fused.wprep(conv), g, bias=b.detach())
    ctx.conv, ctx.ksize, ctx.stride, ctx.xshape = conv, ksize, stride, tuple(x.shape)
    return y

  @staticmethod
  def backward(ctx, dy):
    dy = _c(dy)
    g = ops.geom(ctx.xshape, ctx.ksize, stride=ctx.stride, dil=1, pad=ctx.ksize // 2)
    dx = None
    if ctx.needs_input_grad[0]:
      if ctx.ksize == 3 and ctx.stride == 1:
        dx, _ = fused.conv3x3_c32_dgrad(dy, ctx.conv, g)
      elif ctx.ksize == 5 and ctx.stride == 2:
        dx = fused.conv5x5s2_c32_dgrad(dy, ctx.conv, ctx.xshape[1], ctx.xshape[2])
      else:
        dx, _ = ops.conv_c32(dy, fused.wprep(ctx.conv, 2), ops.geom_transposed(g))
    return dx, dy, ops.channel_sum(dy), None, None, None


def conv_c32_autograd(x, conv, ksize, stride):
  ws = fused.WGRAD_STREAM
  if ws is None or not conv.weight.requires_grad:
    return ConvC32.apply(x, conv.weight, conv.bias, conv, ksize, stride)
  with torch.cuda.stream(ws):
    tok = WgradTokenC32.apply(conv.weight, x.detach(), ksize, stride)
  return ConvC32Split.apply(x, tok, conv.bias, conv, ksize, stride)


class Conv5x5s2First(torch.autograd.Function):
  """downsample[0]: NCHW image -> channels-last features; only weight/bias gradients (the image needs none)."""

  @staticmethod
  def forward(ctx, img, w, b):
    ctx.save_for_backward(img)
    return ops.conv5x5s2_c3(img, w, b)

  @staticmethod
  def backward(ctx, dy):
    (img,) = ctx.saved_tensors
    dy = _c(dy)
    return None, ops.conv5x5s2_c3_wgrad(img, dy), ops.channel_sum(dy)


class CostVolume(torch.autograd.Function):
  @staticmethod
  def forward(ctx, left, right, D):
    return ops.cost_volume(left, right, D)

  @staticmethod
  def backward(ctx, dcost):
    dl, dr = ops.cost_volume_bwd(_c(dcost))
    return dl, dr, None


class Conv3dOutSoftargmin(torch.autograd.Function):
  """conv3d_alone (32->1) + softmax + expectation.  Outputs (cost [B,D,H,W], pred [B,H,W])."""

  @staticmethod
  def forward(ctx, x, w, b):
    cost, pred, fcs = ops.conv3d_out_softargmin(x, w, b, want_cost=True, want_fcs=x.shape[1] > 2)
    ctx.save_for_backward(x, w, cost, pred)
    ctx.mark_non_differentiable(fcs) if fcs is not None else None
    return cost, pred, fcs

  @staticmethod
  def backward(ctx, dcost_out, dpred, _dfcs=None):
    x, w, cost, pred = ctx.saved_tensors
    dcost = ops.softargmin_bwd(cost, pred, _c(dpred), _c(dcost_out))
    dx, dw, db = ops.conv_c32_taps_bwd(x, w, dcost, 27)
    return dx, dw, db


def conv3d_out_softargmin_autograd(x, conv, want_cost):
  cost, pred, fcs = Conv3dOutSoftargmin.apply(x, conv.weight, conv.bias)
  if want_cost and fcs is not None:
    cost._snb_fcs = fcs
  return (cost if want_cost else None), pred


class Upsample(torch.autograd.Function):
  @staticmethod
  def forward(ctx, pred, H, W, mul):
    ctx.hw, ctx.mul = tuple(pred.shape[-2:]), mul
    return ops.upsample_bilinear(pred, H, W, mul)

  @staticmethod
  def backward(ctx, dout):
    return ops.upsample_bilinear_bwd(_c(dout), ctx.hw[0], ctx.hw[1], ctx.mul), None, None, None


class RefineHead(torch.autograd.Function):
  """Upsample + scale + concat + Conv2d(4->32) + BN + LeakyReLU.  Outputs (up [B,H,W], y [B,H,W,32])."""

  @staticmethod
  def forward(ctx, coarse, rgb, w, b, gamma, beta, bn, training):
    up, z, stats = ops.refine_in_conv(coarse, rgb, w, b.detach(), want_stats=training)
    if training:
      scale, shift, mean, invstd = fused.bn_finalize(stats, z.numel() // 32, bn)
    else:
      scale, shift = fused.bn_fold(bn)
      mean, invstd = bn.running_mean, fused.bn_invstd(bn)
    y = ops.bn_apply(z, scale, shift, residual=None, lrelu=True)
    ctx.save_for_backward(coarse, rgb, w, z, scale, shift, mean, invstd)
    ctx.training = training
    return up, y

  @staticmethod
  def backward(ctx, dup, dy):
    coarse, rgb, w, z, scale, shift, mean, invstd = ctx.saved_tensors
    dz, dgamma, dbeta, dbias = ops.bn_lrelu_bwd(z, _c(dy), scale, shift, mean, invstd, ctx.training, lrelu=True)
    dw = ops.refine_in_wgrad(coarse, rgb, dz)
    dcoarse = None
    if ctx.needs_input_grad[0]:
      # d(input channel 0)[q] = sum_tap sum_co dz[q - off(tap)][co] * W[co][0][tap]: a 32->1 "conv" with flipped taps
      wflip = torch.flip(w.detach()[:, 0].reshape(32, 9), dims=(1,)).reshape(1, 32, 3, 3).contiguous()
      taps = ops.conv_c32_taps(dz, wflip, 9)
      d_up = ops.tapsum_refine_out(taps, None, _c(dup), relu=False)          # + gradient arriving directly on `up`
      dcoarse = ops.upsample_bilinear_bwd(d_up, coarse.shape[-2], coarse.shape[-1], float(rgb.shape[-1]) / float(coarse.shape[-1]))
    return dcoarse, None, dw, dbias, dgamma, dbeta, None, None


class WgradTokenRefine(torch.autograd.Function):
  """WgradToken for RefineHead's Conv2d(4 -> 32): see WgradToken."""

  @staticmethod
  def forward(ctx, w, coarse, rgb):
    ctx.save_for_backward(coarse, rgb)
    return fused.wgrad_token((rgb.shape[0], rgb.shape[2], rgb.shape[3], 32), rgb.device)

  @staticmethod
  def backward(ctx, dz):
    coarse, rgb = ctx.saved_tensors
    dz = _c(dz)
    _on_this_stream(coarse, rgb, dz)
    return ops.refine_in_wgrad(coarse, rgb, dz), None, None


class RefineHeadSplit(torch.autograd.Function):
  """RefineHead with the weight gradient split off into WgradTokenRefine."""

  @staticmethod
  def forward(ctx, coarse, rgb, tok, w, b, gamma, beta, bn, training):
    up, z, stats = ops.refine_in_conv(coarse, rgb, w, b.detach(), want_stats=training)
    if training:
      scale, shift, mean, invstd = fused.bn_finalize(stats, z.numel() // 32, bn)
    else:
      scale, shift = fused.bn_fold(bn)
      mean, invstd = bn.running_mean, fused.bn_invstd(bn)
    y = ops.bn_apply(z, scale, shift, residual=None, lrelu=True)
    ctx.save_for_backward(coarse, rgb, w, z, scale, shift, mean, invstd)
    ctx.training = training
    return up, y

  @staticmethod
  def backward(ctx, dup, dy):
    coarse, rgb, w, z, scale, shift, mean, invstd = ctx.saved_tensors
    dz, dgamma, dbeta, dbias = ops.bn_lrelu_bwd(z, _c(dy), scale, shift, mean, invstd, ctx.training, lrelu=True)
    dcoarse = None
    if ctx.needs_input_grad[0]:
      wflip = torch.flip(w[:, 0].reshape(32, 9), dims=(1,)).reshape(1, 32, 3, 3).contiguous()
      taps = ops.conv_c32_taps(dz, wflip, 9)
      d_up = ops.tapsum_refine_out(taps, None, _c(dup), relu=False)
      dcoarse = ops.upsample_bilinear_bwd(d_up, coarse.shape[-2], coarse.shape[-1], float(rgb.shape[-1]) / float(coarse.shape[-1]))
    return dcoarse, None, dz, None, dbias, dgamma, dbeta, None, None


def refine_head_autograd(coarse, rgb, conv, bn, training):
  ws = fused.WGRAD_STREAM
  if ws is None or not conv.weight.requires_grad:
    return RefineHead.apply(coarse, rgb, conv.weight, conv.bias, bn.weight, bn.bias, bn, training)
  with torch.cuda.stream(ws):
    tok = WgradTokenRefine.apply(conv.weight, coarse.detach(), rgb)
  return RefineHeadSplit.apply(coarse, rgb, tok, conv.weight.detach(), conv.bias, bn.weight, bn.bias, bn, training)


class RefineTail(torch.autograd.Function):
  """Conv2d(32->1) + add upsampled disparity + ReLU."""

  @staticmethod
  def forward(ctx, x, w, b, up):
    taps = ops.conv_c32_taps(x, w, 9)
    out = ops.tapsum_refine_out(taps, b, up)
    ctx.save_for_backward(x, w, out)
    return out

  @staticmethod
  def backward(ctx, dout):
    x, w, out = ctx.saved_tensors
    dres = ops.relu_bwd(out, _c(dout))
    dx, dw, db = ops.conv_c32_taps_bwd(x, w, dres, 9)
    return dx, dw, db, dres


def refine_tail_autograd(x, conv, up):
  return RefineTail.apply(x, conv.weight, conv.bias, up)
