"""The reference's import surface for the model: `adaptive_stereo.models.stereo_net.{StereoNet, ...}` must
resolve to the drop-in — through stereonet_b200.install() and through the namespace-package shim directory — while the other
reference modules keep resolving to the reference.  GPU part: the bodies of the reference's own timing / test scripts
(evaluation/stereonet_timing.py:22-41,44-72, test/test_stereo_net.py:25-72) restated with those imports."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "adaptive-stereo-icra-2021_b200")


def _run(code, pythonpath):
  env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath))
  return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=300)


def test_install_resolves_reference_imports(tmp_path):
  # a stand-in for the rest of the reference checkout: adaptive_stereo/utils/marker.py in a namespace package (no __init__.py)
  (tmp_path / "adaptive_stereo" / "utils").mkdir(parents=True)
  (tmp_path / "adaptive_stereo" / "utils" / "marker.py").write_text("WHO = 'reference'\n")
  r = _run("""
      import stereonet_b200 as S
      names = S.install()
      assert names == ["adaptive_stereo.models.stereo_net", "models.stereo_net"], names
      from adaptive_stereo.models.stereo_net import StereoNet, FeatureExtractorNetwork, DisparityRegression, EdgeAwareRefinement, BasicBlock, convbn, convbn_3d
      from models.stereo_net import StereoNet as S2, FeatureExtractorNetwork as F2          # test/test_stereo_net.py:10
      from adaptive_stereo.utils.marker import WHO                                          # the rest of the reference is untouched
      assert StereoNet is S.StereoNet is S2 and FeatureExtractorNetwork is S.FeatureExtractorNetwork is F2 and WHO == "reference"
      f, s = FeatureExtractorNetwork(4), StereoNet(4, 1, 0)                                 # ctor signatures of stereo_net.py:55,138
      assert len(f.state_dict()) == 94 and len(s.state_dict()) == 123
      S.install()                                                                           # idempotent
      print("ok")
      """, [PKG, str(tmp_path)])
  assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_namespace_shim_directory(tmp_path):
  (tmp_path / "adaptive_stereo" / "utils").mkdir(parents=True)
  (tmp_path / "adaptive_stereo" / "utils" / "marker.py").write_text("WHO = 'reference'\n")
  (tmp_path / "adaptive_stereo" / "models").mkdir(parents=True)
  (tmp_path / "adaptive_stereo" / "models" / "stereo_net.py").write_text("raise ImportError('the reference model must be shadowed by the shim')\n")
  (tmp_path / "adaptive_stereo" / "models" / "linear_warping.py").write_text("WHO = 'reference'\n")
  r = _run("""
      from adaptive_stereo.models.stereo_net import StereoNet, FeatureExtractorNetwork, DisparityRegression
      from adaptive_stereo.models.linear_warping import WHO as W1
      from adaptive_stereo.utils.marker import WHO as W2
      import stereonet_b200 as S
      assert StereoNet is S.StereoNet and W1 == W2 == "reference"
      print("ok")
      """, [os.path.join(PKG, "shim"), PKG, str(tmp_path)])
  assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_timing_and_test_scripts_run_on_the_drop_in():
  """evaluation/stereonet_timing.py (k = 4, zeros 1x3x320x1216: inference trials and inference + backprop trials with the
  reference's Monodepth loss on the outputs, Adam lr 1e-4) and test/test_stereo_net.py (k=4 / s=0 and k=3 / s=1 at 320x960,
  with and without the feature-contrast score), few iterations each; loss functions restated by the oracle."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import stereonet_oracle as O
  import stereonet_b200 as S
  S.install()
  from adaptive_stereo.models.stereo_net import StereoNet, FeatureExtractorNetwork
  from models.stereo_net import StereoNet as SN2
  assert SN2 is StereoNet
  # ---- run_inference_trials (stereonet_timing.py:22-41)
  feature_net = FeatureExtractorNetwork(4).cuda(); stereo_net = StereoNet(4, 1, 0).cuda()
  feature_net.eval(); stereo_net.eval()
  with torch.no_grad():
    iml = torch.zeros(1, 3, 320, 1216).cuda(); imr = torch.zeros(1, 3, 320, 1216).cuda()
    for _ in range(3):
      lf, rf = feature_net(iml), feature_net(imr)
      outputs = stereo_net(iml, lf, rf, "l", output_cost_volume=False)
  assert sorted(outputs) == ["pred_disp_l/0", "pred_disp_l/4"] and tuple(outputs["pred_disp_l/0"].shape) == (1, 1, 320, 1216)
  assert torch.isfinite(outputs["pred_disp_l/0"]).all()
  # ---- run_backprop_trials (stereonet_timing.py:44-72): the loss is built from the outputs by plain PyTorch ops
  feature_net = FeatureExtractorNetwork(4).cuda(); stereo_net = StereoNet(4, 1, 0).cuda()
  feature_net.train(); stereo_net.train()
  optimizer = torch.optim.Adam([{"params": stereo_net.parameters()}, {"params": feature_net.parameters()}], lr=1e-4)
  g = torch.Generator().manual_seed(1)
  iml = torch.rand(1, 3, 320, 1216, generator=g).cuda(); imr = torch.rand(1, 3, 320, 1216, generator=g).cuda()
  w0 = stereo_net.conv3d_alone.weight.detach().clone()
  for _ in range(2):
    lf, rf = feature_net(iml), feature_net(imr)
    outputs = stereo_net(iml, lf, rf, "l", output_cost_volume=False)
    loss = O.monodepth_single_loss(iml, imr, outputs["pred_disp_l/0"])        # = monodepth_single_loss of stereonet_timing.py:11-19
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
  assert torch.isfinite(loss) and not torch.equal(w0, stereo_net.conv3d_alone.weight.detach())
  # ---- test/test_stereo_net.py:25-72: (k, input_scale, compute_fcs) on zeros 1x3x320x960 / 1x3x160x480
  images = {0: torch.zeros(1, 3, 320, 960).cuda(), 1: torch.zeros(1, 3, 160, 480).cuda()}
  for k, s, fcs in ((4, 0, False), (3, 1, False), (4, 0, True)):
    feature_net = FeatureExtractorNetwork(k).cuda(); stereo_net = StereoNet(k, 1, s, maxdisp=192).cuda()
    with torch.no_grad():
      for _ in range(2):
        fl, fr = feature_net(images[s]), feature_net(images[s])
        outputs = stereo_net(images[s], fl, fr, "l", output_cost_volume=fcs)
        if fcs:
          score = O.feature_contrast_mean(outputs["cost_volume_l/{}".format(s + k)])      # utils/feature_contrast.py:12-23
          assert torch.isfinite(score).all()
    assert tuple(outputs["pred_disp_l/{}".format(s)].shape) == (1, 1) + tuple(images[s].shape[-2:])
