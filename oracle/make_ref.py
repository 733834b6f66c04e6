"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, made importable where /root/reference does not exist.

The reference is pure Python (no build step), so "building" oracle/_ref means copying, byte for byte, the four modules the path
consists of out of the reference checkout into the git-ignored, gpurun-shipped directory oracle/_ref/ :

    adaptive_stereo/models/stereo_net.py      FeatureExtractorNetwork / StereoNet (the model)
    adaptive_stereo/models/linear_warping.py  LinearWarping (adapt.py:78-86)
    adaptive_stereo/utils/loss_functions.py   monodepth_loss, khamis_robust_loss
    adaptive_stereo/utils/feature_contrast.py feature_contrast_mean

Nothing under oracle/_ref is committed (see .gitignore) and the product package never imports it; bench.py's `--impl reference`
arm and `cpu_baseline` leg time these modules on the host cores (kind = "reference") and fall back to the oracle port
(kind = "port") when the directory is missing.  Run by __graft_entry__.build() in the build container; a SHA-256 manifest of the
copied files is written next to them.  Usage: python oracle/make_ref.py [reference_root]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["adaptive_stereo/models/stereo_net.py", "adaptive_stereo/models/linear_warping.py",
         "adaptive_stereo/utils/loss_functions.py", "adaptive_stereo/utils/feature_contrast.py"]


def make(ref_root="/root/reference"):
  if not os.path.isdir(ref_root):
    return False
  manifest = {}
  for rel in FILES:
    src = os.path.join(ref_root, rel)
    dst = os.path.join(DST, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
  json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
  return True


def load():
  """Import the vendored reference modules (namespace packages, as in the reference).  On a GPU-less host `.cuda()` — hard-coded
  in stereo_net.py:129,177 — is shimmed to identity; returns (stereo_net, linear_warping, loss_functions,
  feature_contrast) or None when oracle/_ref is absent."""
  if not os.path.exists(os.path.join(DST, FILES[0])):
    return None
  import importlib.util
  import torch
  torch.Tensor.cuda = lambda self, *a, **k: self
  torch.nn.Module.cuda = lambda self, *a, **k: self
  mods = []
  for rel in FILES:
    name = "snb_ref_" + os.path.basename(rel)[:-3]
    spec = importlib.util.spec_from_file_location(name, os.path.join(DST, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mods.append(mod)
  return tuple(mods)


if __name__ == "__main__":
  ok = make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
  print("oracle/_ref " + ("written" if ok else "NOT written (no reference checkout here)"))
