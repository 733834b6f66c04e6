"""How far do the adaptation-step gradients of the three fp32-grade convolution back ends (fp16 split, 3xTF32, FFMA) sit from the
reference's golden gradients (k3_b2_sharp)?  Diagnostic for the conditioning of that test."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "adaptive-stereo-icra-2021_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import torch
import stereonet_oracle as O
import stereonet_b200 as S
from stereonet_b200.autograd import fused
from stereonet_b200.adapt import AdaptStepper, make_optimizer
from stereonet_b200.losses import monodepth_single_loss
from test_oracle_golden import CASES, TRAIN_SHARPEN_RATIO, build
DEV = "cuda:0"
for name in ("k3_b2_sharp", "k3_ragged"):
  cfg = CASES[name]
  g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
  res = {}
  for backend in ("h3", "tc3", "ffma"):
    fused.set_conv_backend(backend)
    fsd, _, left, right, _ = build(cfg)
    ssd = O.make_stereo_state(seed=22, sharpen=cfg["sharpen"] * TRAIN_SHARPEN_RATIO)
    f = S.FeatureExtractorNetwork(cfg["k"]).to(DEV); s = S.StereoNet(cfg["k"], 1, cfg["s"]).to(DEV)
    f.load_state_dict(fsd); s.load_state_dict(ssd)
    opt = make_optimizer(f, s, lr=5e-5)
    st = AdaptStepper(f, s, opt, cfg["H"], cfg["W"], clip_grad_norm=False)
    f.train(); s.train()
    l, r = left.to(DEV), right.to(DEV)
    out = st.predict(l, r)
    loss = monodepth_single_loss(l, r, out, cfg["s"])
    opt.zero_grad(); loss.backward()
    worst_norm, worst_entry, worst_cos = 0.0, 0.0, 1.0
    grads = {}
    for tag, net in (("s", s), ("f", f)):
      for n, p in net.named_parameters():
        if p.grad is None: continue
        key = f"grad/{tag}/{n}"
        grads[key] = p.grad.cpu().numpy().ravel().astype(np.float64)
        if key in g.files and not (n.endswith(".0.0.bias") or n == "conv3d_alone.bias"):
          a, b = grads[key], g[key].ravel().astype(np.float64)
          if np.linalg.norm(b) < 1e-6: continue
          cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
          ent = float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))
          if ent > worst_entry: worst_entry, wk = ent, key
          worst_cos = min(worst_cos, cos)
    res[backend] = grads
    print(f"{name} [{backend}] loss {loss.item():.7f} (ref {float(g['train/loss']):.7f}) worst single-entry deviation {worst_entry:.4f} ({wk}) worst cos {worst_cos:.5f}", flush=True)
  for a, b in (("h3", "tc3"), ("h3", "ffma"), ("tc3", "ffma")):
    w = max(float(np.abs(res[a][k] - res[b][k]).max() / (np.abs(res[b][k]).max() + 1e-12)) for k in res[a] if np.linalg.norm(res[b][k]) > 1e-6)
    print(f"  {a} vs {b}: worst single-entry deviation between back ends {w:.4f}")
fused.set_conv_backend("h3")
