"""The A/B switches of the library select between kernels of this library (never a fallback).  Each variant must give the
same results as the default path: the CUDA-core cross-check kernels (SNB200_SMALL_CONV=ffma), launches without programmatic
dependent launch (SNB200_PDL=0) and — selected by the test itself through fused.set_conv_backend, not the environment — the
round-1 kernels with the fp16 split, the 3xTF32 operand split and the fp32 FFMA kernel instead of snb_conv_c32_ws.  The switches are read once
per process, so every variant runs in its own interpreter."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import ops
from stereonet_b200.autograd import fused
from test_gpu_kernels import DEV

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import os, sys
import numpy as np, torch
ROOT = sys.argv[1]
for p in (ROOT, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"), os.path.join(ROOT, "oracle")):
  sys.path.insert(0, p)
import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200 import ops
from stereonet_b200.autograd import fused
if os.environ.get("SNB_TEST_BACKEND"):
  fused.set_conv_backend(os.environ["SNB_TEST_BACKEND"])
fmt = fused._FMT.get(fused.CONV_BACKEND, "ws")
dev = "cuda:0"
k = 3
f = S.FeatureExtractorNetwork(k).to(dev).eval(); s = S.StereoNet(k, 1, 0).to(dev).eval()
f.load_state_dict(O.make_feature_state(k, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=10.0))
l, r, _ = O.make_stereo_pair(2, 100, 260, seed=1000, max_disp_px=40.0)
l, r = l.to(dev), r.to(dev)
with torch.no_grad():
  out = s(l, f(l), f(r), "l", output_cost_volume=True)
  x3 = torch.randn(1, 6, 9, 150, 32, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
  w3 = torch.randn(32, 32, 3, 3, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) * 0.05
  y3, st3 = ops.conv_c32_tc(x3, ops.prep_conv_weights_tc(w3, fmt=fmt), ops.geom(tuple(x3.shape), 3), want_stats=True, fmt=fmt)
torch.cuda.synchronize()
np.savez(sys.argv[2], disp=out["pred_disp_l/0"].cpu().numpy(), coarse=out["pred_disp_l/3"].cpu().numpy(),
         cost=out["cost_volume_l/3"].cpu().numpy(), y3=y3.cpu().numpy(), st3=st3.double().sum(0).cpu().numpy())
'''


def _run(env_extra):
  env = dict(os.environ); env.update(env_extra)
  with tempfile.TemporaryDirectory() as d:
    out = os.path.join(d, "o.npz")
    subprocess.run([sys.executable, "-c", SNIPPET, ROOT, out], check=True, env=env, timeout=600)
    return dict(np.load(out))


@pytest.fixture(scope="module")
def default_run():
  return _run({})


@pytest.mark.parametrize("env", [{"SNB200_SMALL_CONV": "ffma"}, {"SNB200_PDL": "0"}, {"SNB_TEST_BACKEND": "h3"}, {"SNB_TEST_BACKEND": "tc3"},
                                 {"SNB_TEST_BACKEND": "ffma", "SNB200_SMALL_CONV": "ffma", "SNB200_PDL": "0"}])
def test_variant_matches_default(default_run, env):
  got = _run(env)
  exact = env == {"SNB200_PDL": "0"}                     # same kernels, only the launch attribute differs
  for key, ref in default_run.items():
    err = float(np.abs(got[key] - ref).max())
    if exact:
      assert err == 0.0, (key, err)
    elif key in ("disp", "coarse"):
      assert err <= 1e-2, (key, err)                     # the parity tolerance of the path (px)
    else:
      assert err <= 2e-5 * max(1.0, float(np.abs(ref).max())), (key, err)


def test_weight_prep_batch_matches_per_layer_prep():
  """fused.WeightPrepBatch (one launch for every derived weight image of the adaptation step) against the per-layer
  snb_prep_conv_weights_tc calls it replaces, including the polyphase sub-kernels of the 5x5 stride-2 layers."""
  f = S.FeatureExtractorNetwork(3).to(DEV); s = S.StereoNet(3, 1, 0).to(DEV)
  f.load_state_dict(O.make_feature_state(3, 11)); s.load_state_dict(O.make_stereo_state(22, sharpen=10.0))
  batch = fused.WeightPrepBatch([f, s])
  batch.refresh()
  torch.cuda.synchronize()
  n3x3 = n5x5 = 0
  for conv, key, views, _ in batch.items:
    mode = key[1]
    if key[0] == "wtc":
      ref = ops.prep_conv_weights_tc(conv.weight, mode, fmt=key[2])
      ws3d = key[2] == "ws" and conv.weight.dim() == 5            # nine 12 KB images + the 2^-s slot (+ 3 floats of padding)
      n = 9 * 3072 + 1 if ws3d else views[0].numel()
      assert torch.equal(views[0][:n], ref[:n]), (key, tuple(conv.weight.shape))
      slot = 9 * 3072 if ws3d else 3072 + 16
      assert float(views[0][slot]) in [2.0 ** -e for e in range(-20, 41)]    # the 2^-s slot
      n3x3 += 1
    elif key[0] == "wtc_p4":
      ref = ops.prep_conv5x5s2_weights_ws(conv.weight.detach())
      n = 10 * 3072 + 1                                   # ten 12 KB images + the 2^-s slot (the 3 floats after it are padding)
      assert torch.equal(views[0][:n], ref[:n]), key
      assert fused.wprep_p4(conv) is views[0]
      n5x5 += 1
      continue
    else:
      w = conv.weight.detach()
      i = 0
      for a in (0, 1):
        for b in (0, 1):
          sub = w[:, :, a::2, b::2]
          sub = torch.nn.functional.pad(sub, (0, 3 - sub.shape[3], 0, 3 - sub.shape[2])).contiguous()
          # the batched prep scales by max|w| of the whole 5x5 kernel, the per-layer prep by max|sub-kernel|: both are exact
          # power-of-two scalings, so the images may differ but the convolutions they drive must not
          xs = torch.randn(1, 9, 40, 32, device=DEV, generator=torch.Generator(device=DEV).manual_seed(a * 2 + b))
          g3 = ops.geom(tuple(xs.shape), 3)
          y_b, _ = ops.conv_c32_tc(xs, views[i], g3, fmt=key[2])
          y_l, _ = ops.conv_c32_tc(xs, ops.prep_conv_weights_tc(sub, mode, fmt=key[2]), g3, fmt=key[2])
          assert float((y_b - y_l).abs().max()) <= 2e-6 * max(1.0, float(y_l.abs().max())), (key, a, b)
          i += 1
      n5x5 += 1
    # the per-module cache now serves these images
    hit = fused.wprep_tc(conv, mode) if key[0] == "wtc" else fused.wprep_tc_phases(conv, mode)
    assert (hit is views[0]) if key[0] == "wtc" else (hit is views)
  assert n3x3 == 2 * 17 and n5x5 == 2 * 2, (n3x3, n5x5)
