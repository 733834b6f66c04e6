"""Import-surface shim: `adaptive_stereo.models.stereo_net` resolving to the B200 drop-in (stereonet_b200).

Put this directory (`.../adaptive-stereo-icra-2021_b200/shim`) and the package directory (`.../adaptive-stereo-icra-2021_b200`)
on PYTHONPATH *before* the reference checkout.  `adaptive_stereo` and `adaptive_stereo.models` are namespace packages in the
reference (no __init__.py), so everything else — adaptive_stereo.utils.*, adaptive_stereo.datasets.*, models.linear_warping —
keeps resolving to the reference, and the callers that import the model classes (train.py:19-22, adapt.py:50-51,65-75,
evaluate_model.py:52-60, evaluation/stereonet_timing.py:7, evaluation/ood_analysis.py:13, ros/stereo_depth_node.py:129-133)
run unmodified on the CUDA kernels of libsnb200.so.  Equivalent at run time: `import stereonet_b200; stereonet_b200.install()`.
"""
from stereonet_b200.models.stereo_net import (BasicBlock, DisparityRegression, EdgeAwareRefinement, FeatureExtractorNetwork,  # noqa: F401
                                              StereoNet, convbn, convbn_3d)

__all__ = ["FeatureExtractorNetwork", "StereoNet", "DisparityRegression", "EdgeAwareRefinement", "BasicBlock", "convbn", "convbn_3d"]
