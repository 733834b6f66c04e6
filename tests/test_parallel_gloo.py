"""CPU, world_size 2 over gloo: the sharding and gradient all-reduce helpers of stereonet_b200.parallel."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereonet_b200 import parallel
import stereonet_b200 as S


def test_shard_range_partitions_exactly():
  for n in (0, 1, 7, 8, 32, 33):
    for world in (1, 2, 4, 8):
      spans = [parallel.shard_range(n, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [b - a for a, b in spans]
      assert max(sizes) - min(sizes) <= 1


def test_used_parameters_excludes_dead_conv2():
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  used = parallel.used_parameters(s, f)
  assert sum(p.numel() for p in used) == 288066          # 169 250 + 118 816 (SURVEY §7 "Unused parameters")
  assert used[0] is s.filter[0][0][0].weight              # stereo_net first, adapt.py:208-210


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  used = parallel.used_parameters(s, f)
  g = torch.Generator().manual_seed(100 + rank)
  for i, p in enumerate(used):
    if rank == 1 and i % 7 == 0:
      p.grad = None                       # a rank that skipped its update contributes zeros (adapt.py:385)
    else:
      p.grad = torch.randn(p.shape, generator=g)
  n = parallel.allreduce_gradients(used)
  out[rank] = (n, [p.grad.clone() for p in used[:5]] + [used[7].grad.clone()])
  dist.destroy_process_group()


def test_allreduce_gradients_matches_single_process_mean():
  world, port = 2, 29517
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  assert out[0][0] == 288066
  # recompute the expected mean in one process
  torch.manual_seed(0)
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  used = parallel.used_parameters(s, f)
  gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
  exp = []
  for i, p in enumerate(used):
    g0 = torch.randn(p.shape, generator=gens[0])
    g1 = torch.zeros(p.shape) if i % 7 == 0 else torch.randn(p.shape, generator=gens[1])
    exp.append((g0 + g1) / 2)
  for r in range(world):
    got = out[r][1]
    for a, b in zip(got[:5], exp[:5]):
      assert torch.allclose(a, b, atol=1e-7)
    assert torch.allclose(got[5], exp[7], atol=1e-7)


def test_used_bn_buffers_excludes_dead_conv2():
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  bufs = parallel.used_bn_buffers(s, f)
  # stereo_net: 4 filter BNs + conv2d_feature + 6 blocks = 11; feature_net: 6 blocks; two buffers each
  assert len(bufs) == 2 * (11 + 6) and all(b.numel() == 32 for b in bufs)
  assert bufs[0] is s.filter[0][0][1].running_mean and bufs[1] is s.filter[0][0][1].running_var


def _bcast_worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(10 + rank)                                  # different initial weights per rank
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  with torch.no_grad():
    s.filter[1][0][1].running_mean.fill_(float(rank) + 0.5)
  parallel.broadcast_state(s, f, src=0)
  out[rank] = (s.conv3d_alone.weight.detach().clone(), f.conv_alone.bias.detach().clone(), s.filter[1][0][1].running_mean.clone(),
               int(s.filter[1][0][1].num_batches_tracked))
  dist.destroy_process_group()


def test_broadcast_state_makes_replicas_identical():
  world, port = 2, 29518
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_bcast_worker, args=(world, port, out), nprocs=world, join=True)
  for a, b in zip(out[0][:3], out[1][:3]):
    assert torch.equal(a, b)
  assert float(out[1][2][0]) == 0.5


def test_replay_reservoir_follows_algorithm_r():
  """reservoir.ReplayReservoir mirrors utils/stereo_reservoir.py:5-64: fills up, rejects duplicate indices, then replaces with
  probability max_size / i; seeded draws are reproducible."""
  from stereonet_b200.reservoir import ReplayReservoir
  t = torch.zeros(1)
  r = ReplayReservoir(4, seed=3)
  assert all(r.add(t + i, t, t, i) for i in range(4)) and r.size() == 4
  assert r.add(t, t, t, 2) is False and r.i == 5                 # duplicate index: counted as streamed, not stored
  added = sum(bool(r.add(t + i, t, t, i)) for i in range(10, 410))
  assert r.size() == 4 and 5 <= added <= 60                      # expectation ~ 4 * ln(405 / 5) = 17.6
  assert len({it[0] for it in r.buf}) == 4 and r.indices == {it[0] for it in r.buf}
  a = ReplayReservoir(4, seed=7); b = ReplayReservoir(4, seed=7)
  for i in range(50):
    a.add(t + i, t, t, i); b.add(t + i, t, t, i)
  assert [float(a.sample()[0]) for _ in range(8)] == [float(b.sample()[0]) for _ in range(8)]
