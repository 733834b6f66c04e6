"""CPU-only: drop-in contract of the nn.Modules."""
import os

import pytest
import torch

import stereonet_oracle as O
import stereonet_b200 as S


@pytest.mark.parametrize("k", [3, 4])
def test_state_dict_layout_matches_reference_keys(k):
  f, s = S.FeatureExtractorNetwork(k), S.StereoNet(k, 1, 0)
  ref_f, ref_s = O.make_feature_state(k, 1), O.make_stereo_state(2)   # validated strict=True against the reference in gen_golden
  assert list(f.state_dict().keys()) == list(ref_f.keys()) or sorted(f.state_dict()) == sorted(ref_f)
  assert sorted(s.state_dict()) == sorted(ref_s)
  for key, v in f.state_dict().items():
    assert tuple(v.shape) == tuple(ref_f[key].shape) and v.dtype == ref_f[key].dtype, key
  for key, v in s.state_dict().items():
    assert tuple(v.shape) == tuple(ref_s[key].shape) and v.dtype == ref_s[key].dtype, key
  f.load_state_dict(ref_f, strict=True)
  s.load_state_dict(ref_s, strict=True)
  assert len(f.state_dict()) == 92 + 2 * (k - 3) and len(s.state_dict()) == 123
  assert sum(p.numel() for p in s.parameters()) == 225122


def test_parameter_registration_order():
  s = S.StereoNet(3, 1, 0)
  names = [n for n, _ in s.named_parameters()]
  assert names[0] == "filter.0.0.0.weight" and names[16] == "conv3d_alone.weight"
  assert names[18].startswith("edge_aware_refinements.0.conv2d_feature")
  assert names[-2] == "edge_aware_refinements.0.conv2d_out.weight"
  f = S.FeatureExtractorNetwork(3)
  fn = [n for n, _ in f.named_parameters()]
  assert fn[0] == "downsample.0.weight" and fn[-2] == "conv_alone.weight"


@pytest.mark.skipif(not os.path.exists("/root/reference/adaptive_stereo/models/stereo_net.py"),
                    reason="reference checkout only exists in the build container")
def test_seeded_init_and_strict_load_against_reference_classes():
  import importlib.util
  spec = importlib.util.spec_from_file_location("ref_sn", "/root/reference/adaptive_stereo/models/stereo_net.py")
  ref = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(ref)
  torch.manual_seed(123)
  rf, rs = ref.FeatureExtractorNetwork(3), ref.StereoNet(3, 1, 0)
  torch.manual_seed(123)
  f, s = S.FeatureExtractorNetwork(3), S.StereoNet(3, 1, 0)
  for (n1, p1), (n2, p2) in zip(list(rf.state_dict().items()) + list(rs.state_dict().items()),
                                list(f.state_dict().items()) + list(s.state_dict().items())):
    assert n1 == n2 and torch.equal(p1, p2), n1      # same keys, same order, same RNG consumption
  rs.load_state_dict(s.state_dict(), strict=True)
  s.load_state_dict(rs.state_dict(), strict=True)
  assert [n for n, _ in rs.named_parameters()] == [n for n, _ in s.named_parameters()]


@pytest.mark.parametrize("k", [3, 4])
def test_seeded_init_and_layout_against_reference_fingerprint(k):
  """Same check as above against tests/golden/state_layout.json (written by oracle/gen_golden.py from the reference classes),
  so it also runs where the reference checkout does not exist: state_dict keys, order, shapes, dtypes, the parameter
  registration order (adam.pth indices, train.py:136-137) and the values of a seed-123 construction (RNG consumption order)."""
  import json
  import numpy as np
  gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_layout.json")))
  torch.manual_seed(123)
  f, s = S.FeatureExtractorNetwork(k), S.StereoNet(k, 1, 0, maxdisp=192)
  for tag, net in (("feature_net", f), ("stereo_net", s)):
    ref = gold[f"k{k}/{tag}"]
    got = list(net.state_dict().items())
    assert [n for n, _ in got] == [r[0] for r in ref["state_dict"]], tag
    assert [n for n, _ in net.named_parameters()] == ref["parameters"], tag
    for (n, v), (rn, shape, dtype, summ) in zip(got, ref["state_dict"]):
      assert list(v.shape) == shape and str(v.dtype) == dtype, n
      t = v.detach().double().flatten()
      mine = np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().sqrt().item()])
      np.testing.assert_allclose(mine, np.array(summ), rtol=1e-12, atol=0, err_msg=n)


def test_cpu_tensors_fail_loudly():
  f = S.FeatureExtractorNetwork(3)
  with pytest.raises(RuntimeError, match="no CPU path"):
    f(torch.zeros(1, 3, 64, 64))


def test_optimizer_step_invalidates_derived_weight_caches():
  """Fused optimizers update parameters without bumping tensor version counters; the derived-weight caches must still be
  invalidated (global optimizer post-step hook -> cache epoch)."""
  import torch
  from stereonet_b200.autograd import fused
  p = torch.nn.Parameter(torch.randn(4, 4))
  p.grad = torch.randn(4, 4)
  opt = torch.optim.Adam([p], lr=1e-3, fused=True)
  calls = []
  owner = torch.nn.Module()
  fused._cached(owner, "k", [p], lambda: calls.append(1) or len(calls))
  fused._cached(owner, "k", [p], lambda: calls.append(1) or len(calls))
  assert len(calls) == 1
  opt.step()
  fused._cached(owner, "k", [p], lambda: calls.append(1) or len(calls))
  assert len(calls) == 2
