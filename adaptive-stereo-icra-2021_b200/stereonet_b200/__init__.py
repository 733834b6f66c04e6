"""stereonet_b200 — B200-native (sm_100a) StereoNet forward / online-adaptation hot path.

Host-side mirror of adaptive_stereo.models.stereo_net (miloknowles/adaptive-stereo-icra-2021): same nn.Module
constructors, forward() signatures, output-dict keys and state_dict layout; the arithmetic runs in hand-written CUDA
kernels reached through the C-ABI library libsnb200.so (include/snb200.h).  There is no CPU or eager fallback.
"""
from .models.stereo_net import (FeatureExtractorNetwork, StereoNet, EdgeAwareRefinement, DisparityRegression,
                                BasicBlock, convbn, convbn_3d)

__all__ = ["FeatureExtractorNetwork", "StereoNet", "EdgeAwareRefinement", "DisparityRegression", "BasicBlock",
           "convbn", "convbn_3d"]
