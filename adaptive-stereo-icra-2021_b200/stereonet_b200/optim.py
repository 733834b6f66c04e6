"""clip_grad_norm_(stereo_net.parameters(), 1.0) + torch.optim.Adam of the adaptation step (adapt.py:208-210,391-393) as three
kernel launches over one flat gradient bucket (csrc/optim.cu).

The bucket holds the gradients of the USED parameters in optimizer order — stereo_net first, then feature_net (adapt.py:208-210);
BasicBlock.conv2 is constructed but never called (stereo_net.py:40,44-51), never gets a gradient and torch's Adam skips it, so it
has no slot here either.  The same flat buffer is what shared-model data parallelism all-reduces (parallel.py): no per-parameter
pack / unpack.  `state_dict()` / `load_state_dict()` speak torch.optim.Adam's format, so `adam.pth` (train.py:136-137) round-trips.
"""
import ctypes as C

import torch

from . import _cabi, ops
from ._cabi import check


class FusedAdamClip:
  def __init__(self, stereo_net, feature_net, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, clip_norm=1.0):
    self.lr, self.betas, self.eps, self.clip_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), clip_norm
    self.all_params, self.used, self.used_index = [], [], []          # optimizer order; used_index = position in all_params
    self.group_sizes = []
    n_stereo = 0
    for gi, net in enumerate((stereo_net, feature_net)):
      names = list(net.named_parameters())
      self.group_sizes.append(len(names))
      for name, p in names:
        if ".conv2." not in name:
          self.used_index.append(len(self.all_params))
          self.used.append(p)
          if gi == 0:
            n_stereo += p.numel()
        self.all_params.append(p)
    self.n_clip = n_stereo
    self.offsets, off = [], 0
    for p in self.used:
      self.offsets.append(off)
      off += p.numel()
    self.n_total = off
    dev = self.used[0].device
    if dev.type != "cuda":
      raise RuntimeError("stereonet_b200: FusedAdamClip needs CUDA parameters; this build has no CPU path")
    self.flat_grad = torch.zeros(self.n_total, device=dev, dtype=torch.float32)
    self.exp_avg = torch.zeros_like(self.flat_grad)
    self.exp_avg_sq = torch.zeros_like(self.flat_grad)
    self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
    self.ws = torch.zeros(_cabi.lib().snb_adam_clip_workspace_bytes() // 8 + 1, device=dev, dtype=torch.float64)
    self._table, self._table_key = None, None
    self.param_groups = [{"params": list(stereo_net.parameters()), "lr": self.lr}, {"params": list(feature_net.parameters()), "lr": self.lr}]

  # ------------------------------------------------------------------------------------------------------------------
  def _chunk_table(self):
    key = tuple(p.data_ptr() for p in self.used)
    if self._table_key != key:
      rows = []
      for p, off in zip(self.used, self.offsets):
        if not p.is_contiguous() or p.dtype != torch.float32:
          raise RuntimeError("stereonet_b200: FusedAdamClip needs contiguous fp32 parameters")
        n = p.numel()
        for c in range(0, n, 1024):
          rows.append([p.data_ptr() + 4 * c, off + c, min(1024, n - c)])
      self._table = torch.tensor(rows, dtype=torch.int64).to(self.flat_grad.device)
      self._table_key = key
    return self._table

  def zero_grad(self, set_to_none=True):
    for p in self.all_params:
      p.grad = None

  def pack(self):
    """Gather the gradients autograd left on the used parameters into the flat bucket (a parameter without a gradient
    contributes zeros).  Two launches for the 96 used tensors of a k = 3 model; capturable (host tables travel as kernel args)."""
    n = len(self.used)
    src = (C.c_void_p * n)(*[None if p.grad is None else p.grad.data_ptr() for p in self.used])
    for p in self.used:
      if p.grad is not None and not (p.grad.is_contiguous() and p.grad.dtype == torch.float32):
        raise RuntimeError("stereonet_b200: gradients must be contiguous fp32")
    off = (C.c_longlong * n)(*self.offsets)
    cnt = (C.c_longlong * n)(*[p.numel() for p in self.used])
    check(_cabi.lib().snb_multi_gather(src, off, cnt, n, ops._p(self.flat_grad), ops._stream(self.flat_grad)), "snb_multi_gather")
    ops._count((n + 63) // 64)
    return self.flat_grad

  def step(self, grad_scale=1.0, packed=False):
    """clip (stereo_net slice of the bucket) + Adam.  grad_scale = 1 / world after a SUM all-reduce of the bucket.
    packed=True: the bucket is already filled (pack() was called, possibly followed by an all-reduce)."""
    if not packed:
      self.pack()
    tbl = self._chunk_table()
    check(_cabi.lib().snb_adam_clip_step(ops._p(tbl), tbl.shape[0], ops._p(self.flat_grad), ops._p(self.exp_avg), ops._p(self.exp_avg_sq),
                                         ops._p(self.step_t), self.n_total, self.n_clip,
                                         float(self.clip_norm) if self.clip_norm else 0.0, float(grad_scale), float(self.param_groups[0]["lr"]),
                                         self.betas[0], self.betas[1], self.eps, ops._p(self.ws), ops._stream(self.flat_grad)),
          "snb_adam_clip_step")
    ops._count(3)
    from .autograd import fused
    fused.bump_epoch()                         # parameters were rewritten behind torch's version counters

  def grad_norm(self):
    """Total norm of the clipped group as seen by the last step (1-element device tensor, no sync)."""
    nf = _cabi.lib().snb_adam_clip_workspace_bytes() // 4
    return self.ws.view(torch.float32)[nf - 1:nf]

  # ------------------------------------------------------------------------------------------------------------------ checkpoint compat
  def state_dict(self):
    """torch.optim.Adam format: state keyed by the parameter's index in [stereo_net.parameters(), feature_net.parameters()]."""
    state = {}
    t = float(self.step_t.item())
    if t > 0:
      for idx, p, off in zip(self.used_index, self.used, self.offsets):
        n = p.numel()
        state[idx] = {"step": torch.tensor(t), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                      "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
    groups, base = [], 0
    for g, size in zip(self.param_groups, self.group_sizes):
      groups.append({"lr": g["lr"], "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                     "params": list(range(base, base + size))})
      base += size
    return {"state": state, "param_groups": groups}

  def load_state_dict(self, sd):
    self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.step_t.zero_()
    for idx, p, off in zip(self.used_index, self.used, self.offsets):
      st = sd["state"].get(idx, sd["state"].get(str(idx)))
      if st is None:
        continue
      n = p.numel()
      self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1)); self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
      self.step_t.fill_(float(st["step"]))
    for g, sg in zip(self.param_groups, sd["param_groups"]):
      g["lr"] = sg["lr"]
