"""The oracle (oracle/stereonet_oracle.py) against golden vectors produced by the reference module itself
(oracle/gen_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

import stereonet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN_SHARPEN_RATIO = 0.25      # oracle/gen_golden.py: milder conv3d_alone gain for the train-mode step
CASES = {
  "k3_small_dgtw": dict(B=1, H=64, W=128, k=3, s=0, sharpen=1.0, train=False),
  "k3_b2_sharp":   dict(B=2, H=96, W=256, k=3, s=0, sharpen=40.0, train=True),
  "k4_sharp":      dict(B=1, H=64, W=256, k=4, s=0, sharpen=40.0, train=False),
  "k3_ragged":     dict(B=1, H=68, W=120, k=3, s=0, sharpen=40.0, train=True),
  "k3_scale1":     dict(B=1, H=48, W=96,  k=3, s=1, sharpen=40.0, train=False),
}


def build(cfg):
  fsd = O.make_feature_state(cfg["k"], seed=11)
  ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"])
  left, right, gt = O.make_stereo_pair(cfg["B"], cfg["H"], cfg["W"], seed=1000, max_disp_px=min(60.0, cfg["W"] / 4))
  return fsd, ssd, left, right, gt


def summ(t):
  t = t.detach().double().flatten()
  return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().sqrt().item()])


@pytest.mark.parametrize("name", list(CASES))
def test_inputs_regenerate(name):
  """Seeded weights/inputs must regenerate exactly, otherwise the fixtures do not apply."""
  cfg = CASES[name]
  g = np.load(os.path.join(GOLD, name + ".npz"))
  fsd, ssd, left, right, _ = build(cfg)
  np.testing.assert_allclose(summ(left), g["in_left_sum"], rtol=1e-12)
  np.testing.assert_allclose(summ(right), g["in_right_sum"], rtol=1e-12)
  np.testing.assert_allclose(summ(torch.cat([v.flatten().float() for v in fsd.values()])), g["w_feat_sum"], rtol=1e-12)
  np.testing.assert_allclose(summ(torch.cat([v.flatten().float() for v in ssd.values()])), g["w_stereo_sum"], rtol=1e-12)


@pytest.mark.parametrize("name", list(CASES))
def test_eval_forward_matches_reference(name):
  cfg = CASES[name]
  g = np.load(os.path.join(GOLD, name + ".npz"))
  fsd, ssd, left, right, gt = build(cfg)
  ex = {}
  with torch.no_grad():
    o = O.predict_disparity_left(fsd, ssd, left, right, cfg["k"], cfg["s"], training=False, extras=ex)
  np.testing.assert_allclose(ex["left_features"].numpy(), g["eval/left_features"], rtol=0, atol=2e-5)
  np.testing.assert_allclose(ex["right_features"].numpy(), g["eval/right_features"], rtol=0, atol=2e-5)
  for key, v in o.items():
    ref = g["eval/" + key]
    tol = 1e-3 if key.startswith("pred_disp") else 2e-4
    np.testing.assert_allclose(v.numpy(), ref, rtol=0, atol=tol, err_msg=key)
  s, k = cfg["s"], cfg["k"]
  assert abs(O.feature_contrast_mean(o[f"cost_volume_l/{s + k}"]).mean().item() - float(g["eval/fcs"])) < 1e-4
  assert abs(O.epe(o[f"pred_disp_l/{s}"], gt).item() - float(g["eval/epe"])) < 1e-3
  # the numpy literal-loop restatement of the cost volume is bit-identical to the differentiable one
  D = (192 + 1) // 2 ** (s + k)
  cv = O.cost_volume_numpy(ex["left_features"].numpy(), ex["right_features"].numpy(), D)
  assert np.array_equal(cv, ex["raw_cost_volume"].numpy())


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["train"]])
def test_adapt_step_matches_reference(name):
  cfg = CASES[name]
  g = np.load(os.path.join(GOLD, name + ".npz"))
  fsd, ssd, left, right, _ = build(cfg)
  ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"] * TRAIN_SHARPEN_RATIO)
  fsd, ssd = O.clone_state(fsd, True), O.clone_state(ssd, True)
  adam = {}
  loss, outputs, grads = O.adapt_step(fsd, ssd, left, right, cfg["k"], adam, lr=5e-5, input_scale=cfg["s"])
  assert abs(loss.item() - float(g["train/loss"])) < 2e-6
  for key, v in outputs.items():
    np.testing.assert_allclose(v.detach().numpy(), g["train/" + key], rtol=0, atol=2e-3, err_msg=key)
  # unclipped gradients were stored; recompute them from the clipped ones via the stored norm
  gn = float(g["train/grad_norm_stereo"])
  coef = min(1.0, 1.0 / (gn + 1e-6))
  checked = 0
  for key in g.files:
    if key.startswith("grad_sum/"):
      _, tag, n = key.split("/", 2)
      got = grads[(tag, n)] / (coef if tag == "s" else 1.0)
      ref = g[key]
      assert abs(summ(got)[2] - ref[2]) <= 2e-3 * ref[2] + 1e-7, (key, summ(got), ref)
      checked += 1
    elif key.startswith("grad_none/"):
      _, tag, n = key.split("/", 2)
      assert O.unused_param(n)
    elif key.startswith("grad/"):
      _, tag, n = key.split("/", 2)
      got = (grads[(tag, n)] / (coef if tag == "s" else 1.0)).numpy()
      ref = g[key]
      scale = np.abs(ref).max() + 1e-12
      assert np.abs(got - ref).max() <= 2e-3 * scale + 1e-7, key
  assert checked > 40
  for key in g.files:
    if key.startswith("post/"):
      _, tag, n = key.split("/", 2)
      sd = ssd if tag == "s" else fsd
      np.testing.assert_allclose(sd[n].detach().numpy(), g[key], rtol=1e-4, atol=1e-6, err_msg=key)
    elif key.startswith("post_sum/"):
      _, tag, n = key.split("/", 2)
      sd = ssd if tag == "s" else fsd
      np.testing.assert_allclose(summ(sd[n]), g[key], rtol=1e-5, atol=1e-7, err_msg=key)
