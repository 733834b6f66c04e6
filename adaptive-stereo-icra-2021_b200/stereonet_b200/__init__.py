"""stereonet_b200 — B200-native (sm_100a) StereoNet forward / online-adaptation hot path.

Host-side mirror of adaptive_stereo.models.stereo_net (miloknowles/adaptive-stereo-icra-2021): same nn.Module
constructors, forward() signatures, output-dict keys and state_dict layout; the arithmetic runs in hand-written CUDA
kernels reached through the C-ABI library libsnb200.so (include/snb200.h).  There is no CPU or eager fallback.
"""
from .models.stereo_net import (FeatureExtractorNetwork, StereoNet, EdgeAwareRefinement, DisparityRegression,
                                BasicBlock, convbn, convbn_3d)

__all__ = ["FeatureExtractorNetwork", "StereoNet", "EdgeAwareRefinement", "DisparityRegression", "BasicBlock",
           "convbn", "convbn_3d", "install"]


def install(short_names=True):
  """Make `adaptive_stereo.models.stereo_net` (and, with short_names, `models.stereo_net` as test/test_stereo_net.py:10 imports it)
  resolve to this drop-in in the running interpreter, leaving every other reference module (adaptive_stereo.utils.*, .datasets.*,
  models.linear_warping, ...) where it is.  Call it before the reference's entry points import the model; idempotent.  The static
  alternative is the namespace-package shim directory `adaptive-stereo-icra-2021_b200/shim`."""
  import importlib
  import sys
  import types
  from .models import stereo_net as impl

  def ensure_pkg(name):
    try:
      return importlib.import_module(name)
    except ImportError:
      pkg = types.ModuleType(name)
      pkg.__path__ = []                      # a package without files of its own
      sys.modules[name] = pkg
      parent, _, leaf = name.rpartition(".")
      if parent:
        setattr(ensure_pkg(parent), leaf, pkg)
      return pkg

  installed = []
  for full in (["adaptive_stereo.models.stereo_net"] + (["models.stereo_net"] if short_names else [])):
    parent, _, leaf = full.rpartition(".")
    mod = types.ModuleType(full)
    mod.__doc__ = "stereonet_b200 drop-in for " + full
    for n in ("FeatureExtractorNetwork", "StereoNet", "DisparityRegression", "EdgeAwareRefinement", "BasicBlock", "convbn", "convbn_3d"):
      setattr(mod, n, getattr(impl, n))
    mod.__all__ = [n for n in dir(mod) if not n.startswith("_")]
    sys.modules[full] = mod
    setattr(ensure_pkg(parent), leaf, mod)
    installed.append(full)
  return installed
