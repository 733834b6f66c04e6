"""Device-resident validation / evaluation: the two periodic, bursty consumers of the hot
path in the reference's adaptation loop, batched and without per-pair host syncs.

  validate_ovs     StateMachine.validate, adapt.py:122-142: re-score every pair of the online validation set (<= 16 reservoir
                   pairs) with the current weights — eval mode, no grad, single-image Monodepth loss per pair.  The reference
                   runs the pairs one at a time and calls .item() on each loss; here the pairs go through the model in
                   batches (eval-mode BatchNorm uses running statistics, so samples do not interact) and the per-pair losses
                   come back as ONE device tensor from snb_photo_loss.
  evaluate_batch   the per-batch body of train.evaluate, train.py:89-110: EPE over gt > 0, D1-all at 2/3/4/5 px and the mean
                   feature-contrast score, as device scalars (snb_eval_metrics + snb_feature_contrast).
  evaluate         the loop of train.py:74-126 over an iterable of (left, right, gt) batches; ONE host sync at the end.
"""
import torch

from . import ops
from .losses import feature_contrast_mean

D1_THRESHOLDS = (2, 3, 4, 5)        # train.py:106


def _predict(feature_net, stereo_net, left, right):
  fl, fr = feature_net(left), feature_net(right)                                # adapt.py:72 / train.py:20-21
  return stereo_net(left, fl, fr, "l", output_cost_volume=True)


def validate_ovs(feature_net, stereo_net, lefts, rights, chunk=8):
  """lefts/rights: [n,3,H,W] tensors (or lists of [1,3,H,W] / [3,H,W]) of the OVS pairs.  Returns a float32 device tensor
  [n] with the Monodepth single-image loss of every pair under the current weights (adapt.py:136-139).  The networks are put
  in eval mode for the call and back in train mode afterwards, as the reference does (adapt.py:129-130,141-142)."""
  if isinstance(lefts, (list, tuple)):
    lefts = torch.cat([t if t.dim() == 4 else t.unsqueeze(0) for t in lefts], 0)
    rights = torch.cat([t if t.dim() == 4 else t.unsqueeze(0) for t in rights], 0)
  s = stereo_net.input_scale
  feature_net.eval(); stereo_net.eval()
  out = []
  try:
    with torch.no_grad():
      for i in range(0, lefts.shape[0], chunk):
        l, r = lefts[i:i + chunk].contiguous(), rights[i:i + chunk].contiguous()
        pred = _predict(feature_net, stereo_net, l, r)["pred_disp_l/{}".format(s)]
        loss, _ = ops.photo_loss(l, r, pred.squeeze(1).contiguous(), 1e-3)
        out.append(loss[1:])
  finally:
    feature_net.train(); stereo_net.train()
  return torch.cat(out)


def evaluate_batch(feature_net, stereo_net, left, right, gt_disp):
  """One iteration of train.evaluate (train.py:89-110) for a batch; the networks must already be in eval mode.  Returns a
  dict of 0-dim device tensors with the reference's metric names; nothing is synchronised."""
  s, k = stereo_net.input_scale, stereo_net.k
  with torch.no_grad():
    outputs = _predict(feature_net, stereo_net, left, right)
    pred = outputs["pred_disp_l/{}".format(s)]
    sums = ops.eval_metrics(pred.contiguous(), gt_disp.reshape(pred.shape).contiguous()).sum(0)   # pooled over the batch
    fcs = feature_contrast_mean(outputs["cost_volume_l/{}".format(s + k)]).mean()
  m = {"EPE": sums[0] / sums[1], "FCS": fcs}
  for i, t in enumerate(D1_THRESHOLDS):
    m["D1_all_{}px".format(t)] = sums[2 + i] / sums[1]
  return m


def evaluate(feature_net, stereo_net, batches):
  """train.evaluate (train.py:74-126): mean of the per-batch metrics over `batches` = iterable of (left, right, gt_disp).
  Returns python floats (one device->host copy for the whole evaluation instead of seven .item() calls per batch)."""
  feature_net.eval(); stereo_net.eval()
  rows, names = [], None
  try:
    for left, right, gt in batches:
      m = evaluate_batch(feature_net, stereo_net, left, right, gt)
      names = list(m.keys())
      rows.append(torch.stack([m[n] for n in names]))
  finally:
    feature_net.train(); stereo_net.train()                                    # train.py:123-124
  if not rows:
    return {}
  mean = torch.stack(rows).mean(0).cpu()
  return {n: float(v) for n, v in zip(names, mean)}
