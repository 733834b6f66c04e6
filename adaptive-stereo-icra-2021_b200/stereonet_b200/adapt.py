"""Host-side mirror of one iteration of the reference's online-adaptation loop (adapt.py:304-396, NONSTOP/ER path):
train-mode forward with the cost volume, Monodepth photometric loss, backward, clip_grad_norm_ on the stereo_net
parameters only, Adam step.  The state machine / OVS / logging of adapt.py stay in the caller (out of scope)."""
import torch
import torch.nn as nn

from .losses import LinearWarping, feature_contrast_mean, khamis_robust_loss, monodepth_single_loss


def make_optimizer(feature_net, stereo_net, lr=5e-5):
  """adapt.py:208-210: two parameter groups, stereo_net first."""
  return torch.optim.Adam([{"params": stereo_net.parameters()}, {"params": feature_net.parameters()}], lr=lr)


class AdaptStepper:
  def __init__(self, feature_net, stereo_net, optimizer, height, width, clip_grad_norm=True, er_loss_weight=0.05):
    self.feature_net, self.stereo_net, self.optimizer = feature_net, stereo_net, optimizer
    self.clip, self.er_loss_weight = clip_grad_norm, er_loss_weight
    dev = next(stereo_net.parameters()).device
    self.warper = LinearWarping(height, width, dev)

  def predict(self, left, right):
    fl, fr = self.feature_net(left), self.feature_net(right)                       # adapt.py:72
    return self.stereo_net(left, fl, fr, "l", output_cost_volume=True)             # adapt.py:73

  def step(self, left, right, replay=None, sync_grads=None):
    """One gradient update (adapt.py:313-314,328-337,381-394).  `replay` = (left, right, gt_disp) adds the
    experience-replay term (adapt.py:339-349).  `sync_grads` is called between backward and clip (DP all-reduce)."""
    s = self.stereo_net.input_scale
    self.feature_net.train(); self.stereo_net.train()
    outputs = self.predict(left, right)
    loss = monodepth_single_loss(left, right, outputs, self.warper, s)
    if replay is not None:
      out_er = self.predict(replay[0], replay[1])
      loss = loss + self.er_loss_weight * khamis_robust_loss(out_er["pred_disp_l/{}".format(s)], replay[2])
    fcs = feature_contrast_mean(outputs["cost_volume_l/{}".format(s + self.stereo_net.k)]).mean()
    self.optimizer.zero_grad()
    loss.backward()
    if sync_grads is not None:
      sync_grads()
    if self.clip:
      nn.utils.clip_grad_norm_(self.stereo_net.parameters(), 1.0)                 # adapt.py:391-392
    self.optimizer.step()
    return loss.detach(), fcs, outputs
