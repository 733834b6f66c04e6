"""Host-side mirror of one iteration of the reference's online-adaptation loop (adapt.py:304-396, NONSTOP/ER path):
train-mode forward with the cost volume, Monodepth photometric loss, backward, clip_grad_norm_ on the stereo_net
parameters only, Adam step.  The state machine / OVS / logging of adapt.py stay in the caller (out of scope)."""
import torch
import torch.nn as nn

from .losses import feature_contrast_mean, khamis_robust_loss, monodepth_single_loss


def make_optimizer(feature_net, stereo_net, lr=5e-5, capturable=False, fused=False):
  """adapt.py:208-210: two parameter groups, stereo_net first.  fused=True: optim.FusedAdamClip — clip_grad_norm_ on the stereo_net
  group + Adam as three launches of this library over one flat gradient bucket (always capturable; checkpoint-compatible with
  torch.optim.Adam).  Otherwise torch.optim.Adam as in the reference; capturable=True keeps its step counters on the device
  (needed by AdaptStepper(use_graph=True))."""
  if fused:
    from .optim import FusedAdamClip
    return FusedAdamClip(stereo_net, feature_net, lr=lr)
  # capturable: step counters live on the device; fused: one multi-tensor kernel per group instead of ~220 tiny
  # per-parameter bias-correction kernels (same update rule, torch.optim.Adam defaults)
  return torch.optim.Adam([{"params": stereo_net.parameters()}, {"params": feature_net.parameters()}], lr=lr,
                          capturable=capturable, fused=True if capturable else None)


class AdaptStepper:
  """use_graph=True captures the WHOLE update (forward, loss, backward, [gradient all-reduce,] clip, Adam) in one CUDA
  graph per input shape and replays it per frame: at batch 1 the step is ~700 launches of 3-100 us kernels, so eager
  execution is bound by host launch latency.  Requires an optimizer built with capturable=True.  The experience-replay term
  (`replay=(left, right, gt)`) is captured too (static replay buffers, fused Khamis loss)."""

  def __init__(self, feature_net, stereo_net, optimizer, height=None, width=None, clip_grad_norm=True, er_loss_weight=0.05,
               use_graph=False, batched_replay=False, two_streams=True):
    self.feature_net, self.stereo_net, self.optimizer = feature_net, stereo_net, optimizer
    self.clip, self.er_loss_weight = clip_grad_norm, er_loss_weight
    self.use_graph = use_graph              # (height / width are accepted for call-site compatibility with the reference's warper)
    # run the replay sample through the model in the SAME batch-2 forward/backward as the stream
    # frame instead of a second full pass (adapt.py:339-349).  Off by default: train-mode BatchNorm then normalises with the
    # statistics of both samples together, which is NOT what the reference's two batch-1 passes compute.
    self.batched_replay = batched_replay
    self.two_streams, self._side, self._wstream, self._acc_stream, self._acc_nodes = two_streams, None, None, None, None
    self._rstream, self._rside = None, None
    self._graphs = {}
    self._wprep = None                      # fused.WeightPrepBatch: all derived weight images in one launch per step
    self.launches_per_step = None
    from .optim import FusedAdamClip
    self.fused_opt = isinstance(optimizer, FusedAdamClip)
    if self.fused_opt:
      optimizer.clip_norm = 1.0 if clip_grad_norm else None
    elif use_graph and not all(g.get("capturable", False) for g in optimizer.param_groups):
      raise RuntimeError("AdaptStepper(use_graph=True) needs make_optimizer(..., capturable=True) or fused=True")

  def predict(self, left, right):
    if not (self.two_streams and left.is_cuda and torch.is_grad_enabled()):
      fl, fr = self.feature_net(left), self.feature_net(right)                     # adapt.py:72
      return self.stereo_net(left, fl, fr, "l", output_cost_volume=True)           # adapt.py:73
    # The two feature passes (and, through autograd, their backward passes) are independent chains of small, latency-bound
    # kernels on a fraction of the SMs: the right-image pass runs on a second stream next to the left one.  Both passes update
    # the same BatchNorm buffers, left first in the reference, so the right pass defers its running-statistics updates and
    # they are applied on the main stream after the join.
    from .autograd import fused
    main = torch.cuda.current_stream(left.device)
    if self._side is None:
      self._side = torch.cuda.Stream(device=left.device)
    side = self._side
    side.wait_stream(main)
    fl = self.feature_net(left)
    outer = fused.DEFER_BN                     # not None when this whole pass already defers (the replay pass, below)
    pending = outer if outer is not None else []
    with torch.cuda.stream(side):
      fused.DEFER_BN = pending
      try:
        fr = self.feature_net(right)
      finally:
        fused.DEFER_BN = outer
    main.wait_stream(side)
    if not torch.cuda.is_current_stream_capturing():
      fr.record_stream(main)
    if outer is None:
      fused.flush_deferred_bn(pending)
    return self.stereo_net(left, fl, fr, "l", output_cost_volume=True)

  def step(self, left, right, replay=None, sync_grads=None, dp_params=None, dp_group=None, dp_bucket=None):
    """One gradient update (adapt.py:313-314,328-337,381-394).  `replay` = (left, right, gt_disp) adds the
    experience-replay term (adapt.py:339-349).  Shared-model data parallelism:
      * `dp_bucket` (parallel.DPBucket, with an optim.FusedAdamClip optimizer): gradients are gathered into the flat bucket by one
        kernel, the bucket (gradients + BatchNorm running statistics) is SUM-all-reduced by ONE NCCL call, the statistics are
        averaged in place and the fused clip + Adam kernel applies 1 / world — no per-parameter pack / unpack;
      * `dp_params` (+ `dp_group`) with a torch optimizer: pack -> all-reduce -> unpack (round-1 path);
      * `sync_grads`: a callable run between backward and the update (eager path only).
    With use_graph=True the step is graph(fwd + loss + bwd + gather) -> eager all-reduce -> graph(clip + Adam)."""
    if self.use_graph and sync_grads is None:
      return self._step_graph(left, right, dp_params, dp_group, replay, dp_bucket)
    if dp_params is not None and sync_grads is None and dp_bucket is None:
      from . import parallel
      sync_grads = lambda: parallel.allreduce_gradients(dp_params, dp_group)
    out = self._fwd_bwd(left, right, replay)
    if dp_bucket is not None:
      self.optimizer.pack()
      self._update(grad_scale=dp_bucket.allreduce(dp_group), packed=True)
      return out
    if sync_grads is not None:
      sync_grads()
    self._update()
    return out

  def _fwd_bwd(self, left, right, replay):
    s = self.stereo_net.input_scale
    self.feature_net.train(); self.stereo_net.train()
    self._refresh_weights()
    if replay is not None and self.batched_replay:
      return self._fwd_bwd_batched(left, right, replay)
    from .autograd import fused
    if self.two_streams and left.is_cuda:
      if self._wstream is None:
        self._wstream = torch.cuda.Stream(device=left.device)
        # Gradient accumulation gets a stream of its own.  autograd runs a parameter's AccumulateGrad node on the stream that was
        # current when the node was created and makes THAT stream wait for the producer of every incoming gradient: with the
        # nodes on the main stream, each weight gradient finishing on the side streams would stall the main chain.  The nodes
        # are created here, under the accumulation stream, and kept alive (autograd holds them weakly).
        self._acc_stream = torch.cuda.Stream(device=left.device)
        with torch.cuda.stream(self._acc_stream):
          self._acc_nodes = [p.view_as(p).grad_fn.next_functions[0][0]
                             for net in (self.stereo_net, self.feature_net) for p in net.parameters() if p.requires_grad]
      fused.WGRAD_STREAM = self._wstream                                            # weight gradients next to the data path
      fused.wgrad_token((1,), left.device)                                          # (creates the token buffer outside any capture)
    try:
      if replay is not None and self.two_streams and left.is_cuda:
        # The replay pass (adapt.py:339-349) is independent of the stream pass until the two losses are added: it runs on its own
        # stream (its feature passes side by side on two more), with ALL its BatchNorm running-statistics updates deferred and
        # applied after the stream pass's — the reference's order.
        main = torch.cuda.current_stream(left.device)
        if self._rstream is None:
          self._rstream = torch.cuda.Stream(device=left.device)
          self._rside = torch.cuda.Stream(device=left.device)
        self._rstream.wait_stream(main)
        outputs = self.predict(left, right)
        loss = monodepth_single_loss(left, right, outputs, s)                      # adapt.py:328-337 (snb_photo_loss)
        pend_r = []
        side_main, self._side = self._side, self._rside
        with torch.cuda.stream(self._rstream):
          fused.DEFER_BN = pend_r
          try:
            out_er = self.predict(replay[0], replay[1])
            loss_er = khamis_robust_loss(out_er["pred_disp_l/{}".format(s)], replay[2])
          finally:
            fused.DEFER_BN = None
            self._side = side_main
        main.wait_stream(self._rstream)
        if not torch.cuda.is_current_stream_capturing():
          loss_er.record_stream(main)
        fused.flush_deferred_bn(pend_r)
        loss = loss + self.er_loss_weight * loss_er
      else:
        outputs = self.predict(left, right)
        loss = monodepth_single_loss(left, right, outputs, s)                      # adapt.py:328-337 (snb_photo_loss)
        if replay is not None:
          out_er = self.predict(replay[0], replay[1])                               # adapt.py:339-349: a second full pass
          loss = loss + self.er_loss_weight * khamis_robust_loss(out_er["pred_disp_l/{}".format(s)], replay[2])
    finally:
      fused.WGRAD_STREAM = None
    fcs = feature_contrast_mean(outputs["cost_volume_l/{}".format(s + self.stereo_net.k)]).mean()
    self.optimizer.zero_grad()
    loss.backward()
    return loss.detach(), fcs, outputs

  def _refresh_weights(self):
    from .autograd import fused
    if fused.CONV_BACKEND == "ffma" or not torch.is_grad_enabled():
      return
    if self._wprep is None or not self._wprep.valid():
      self._wprep = fused.WeightPrepBatch([self.feature_net, self.stereo_net])
    self._wprep.refresh()

  def _fwd_bwd_batched(self, left, right, replay):
    """Row f3: one batch-(n+m) pass for the n stream frames and the m replay samples; Monodepth loss on the first n
    predictions, Khamis robust loss on the rest (snb_khamis_loss), same 1 : er_loss_weight mix as adapt.py:383-388."""
    s = self.stereo_net.input_scale
    n = left.shape[0]
    outputs = self.predict(torch.cat([left, replay[0]], 0), torch.cat([right, replay[1]], 0))
    key = "pred_disp_l/{}".format(s)
    head = {k: v[:n] for k, v in outputs.items()}
    loss = monodepth_single_loss(left, right, head, s)
    loss = loss + self.er_loss_weight * khamis_robust_loss(outputs[key][n:], replay[2])
    fcs = feature_contrast_mean(head["cost_volume_l/{}".format(s + self.stereo_net.k)]).mean()
    self.optimizer.zero_grad()
    loss.backward()
    return loss.detach(), fcs, head

  def _update(self, grad_scale=1.0, packed=False):
    if self.fused_opt:
      self.optimizer.step(grad_scale=grad_scale, packed=packed)                   # clip (adapt.py:391-392) + Adam (:393), 3 launches
      return
    if self.clip:
      nn.utils.clip_grad_norm_(self.stereo_net.parameters(), 1.0)                 # adapt.py:391-392
    self.optimizer.step()

  # ---------------------------------------------------------------------------------------------- CUDA-graph path
  def _state_tensors(self):
    ts = [p for net in (self.stereo_net, self.feature_net) for p in net.parameters()]
    ts += [b for net in (self.stereo_net, self.feature_net) for b in net.buffers()]
    if self.fused_opt:
      ts += [self.optimizer.exp_avg, self.optimizer.exp_avg_sq, self.optimizer.step_t]
    else:
      for st in self.optimizer.state.values():
        ts += [v for v in st.values() if isinstance(v, torch.Tensor)]
    return ts

  def _capture(self, left, right, dp_params, dp_group, replay=None, dp_bucket=None):
    import torch.distributed as dist
    from . import ops, parallel
    from .autograd import fused
    self.feature_net.train(); self.stereo_net.train()
    dev = left.device
    sl, sr = left.clone(), right.clone()
    srep = None if replay is None else tuple(t.clone() for t in replay)        # static replay buffers (left, right, gt)
    world = dist.get_world_size(dp_group) if (dp_params is not None or dp_bucket is not None) else 1
    fresh = (not self.fused_opt) and len(self.optimizer.state) == 0
    snap = [(t, t.detach().clone()) for t in self._state_tensors()]
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
      # One eager warm-up step allocates the lazily-created state (Adam moments and step counters, derived-weight
      # caches, allocator pools); model and optimizer state are restored afterwards so that capturing does not advance
      # the adaptation.
      self._fwd_bwd(sl, sr, srep)
      if dp_bucket is not None:
        self.optimizer.pack()
        self._update(grad_scale=dp_bucket.allreduce(dp_group), packed=True)
      else:
        if dp_params is not None:
          parallel.allreduce_gradients(dp_params, dp_group)
        self._update()
      self.optimizer.zero_grad(set_to_none=True)
      with torch.no_grad():
        for t, saved in snap:                 # in place: the graphs will hold these addresses
          t.copy_(saved)
        if fresh:
          for st in self.optimizer.state.values():
            for v in st.values():
              if isinstance(v, torch.Tensor):
                v.zero_()
    cur.wait_stream(side)
    fused.bump_epoch()
    n0 = ops.LAUNCHES
    g1 = torch.cuda.CUDAGraph()
    flat = None
    with torch.cuda.graph(g1):
      out = self._fwd_bwd(sl, sr, srep)
      if dp_bucket is not None:
        self.optimizer.pack()
      elif dp_params is not None:
        flat = parallel.pack_gradients(dp_params)
      else:
        self._update()
    g2 = None
    if dp_bucket is not None:
      g2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g2, pool=g1.pool()):
        with torch.no_grad():
          dp_bucket.stats.mul_(1.0 / world)
        self._update(grad_scale=1.0 / world, packed=True)
      flat = dp_bucket.flat
    elif dp_params is not None:
      g2 = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g2, pool=g1.pool()):
        with torch.no_grad():
          parallel.unpack_gradients(dp_params, flat, world)
        self._update()
    self.launches_per_step = ops.LAUNCHES - n0          # library kernels inside one replay
    return dict(g1=g1, g2=g2, flat=flat, left=sl, right=sr, replay=srep, out=out)

  def _step_graph(self, left, right, dp_params, dp_group, replay=None, dp_bucket=None):
    import torch.distributed as dist
    from .autograd import fused
    key = (tuple(left.shape), str(left.device), None if dp_params is None else len(dp_params), dp_bucket is not None,
           None if replay is None else tuple(tuple(t.shape) for t in replay), self.batched_replay)
    e = self._graphs.get(key)
    if e is None:
      e = self._graphs[key] = self._capture(left, right, dp_params, dp_group, replay, dp_bucket)
    e["left"].copy_(left, non_blocking=True)
    e["right"].copy_(right, non_blocking=True)
    if replay is not None:
      for dst, src in zip(e["replay"], replay):
        dst.copy_(src, non_blocking=True)
    e["g1"].replay()
    if e["g2"] is not None:
      dist.all_reduce(e["flat"], op=dist.ReduceOp.SUM, group=dp_group)      # eager, between the two graphs
      e["g2"].replay()
    fused.bump_epoch()                      # the replay rewrote the parameters behind torch's version counters
    return e["out"]
