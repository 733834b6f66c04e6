"""Multi-GPU partitioning of the hot path: batched inference and independent adaptation streams are
embarrassingly parallel — one process per GPU, contiguous shards, no data-path collective.  The only exchange step in
scope is the gradient all-reduce of shared-model multi-stream adaptation (flat fp32 bucket of the USED parameters)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
  """Contiguous balanced shard [lo, hi) of n_items for `rank` of `world` (first n_items % world ranks get one more)."""
  if not (0 <= rank < world):
    raise ValueError("rank out of range")
  base, rem = divmod(n_items, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def used_parameters(stereo_net, feature_net):
  """Parameters that ever receive a gradient, in optimizer order (stereo_net first, adapt.py:208-210).
  BasicBlock.conv2 is constructed but never called (stereo_net.py:40,44-51) -> excluded, as Adam skips grad=None."""
  out = []
  for net in (stereo_net, feature_net):
    for name, p in net.named_parameters():
      if ".conv2." not in name:
        out.append(p)
  return out


def pack_gradients(params, out=None):
  """Flat fp32 bucket of the gradients of `params` (zeros where a parameter has no gradient).  `out` reuses a buffer."""
  flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
  if out is not None:
    out.copy_(flat)
    return out
  return flat


def unpack_gradients(params, flat, world):
  """p.grad <- flat / world, parameter by parameter (creates the gradient where a rank had none)."""
  flat = flat / world
  off = 0
  for p in params:
    n = p.numel()
    g = flat[off:off + n].view_as(p)
    if p.grad is None:
      p.grad = g.clone()
    else:
      p.grad.copy_(g)
    off += n


def allreduce_gradients(params, group=None):
  """One flat-bucket SUM all-reduce then divide by world size (classic DP; 288 066 floats = 1.15 MB for k=3).
  A rank whose frame went to the validation set contributes zeros so the collective stays matched (adapt.py:385)."""
  world = dist.get_world_size(group)
  flat = pack_gradients(params)
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  unpack_gradients(params, flat, world)
  return flat.numel()
