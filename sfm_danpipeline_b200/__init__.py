"""sfm_danpipeline_b200 -- B200-native all-pairs descriptor matching (the hot path of iTree3DMap).

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of include/sfm_match.h),
``matcher`` (host-side mirror of StructFromMotion::getMatching), ``distributed`` (pair sharding
across ranks + NCCL broadcast/gather) and ``synth`` (seeded descriptor sets for tests/bench).
"""
from ._lib import (BINARY_AUTO, BINARY_POPC, BINARY_TENSOR, DMATCH_DTYPE, FLOAT_AUTO, FLOAT_EXACT, FLOAT_TENSOR, NORM_HAMMING,  # noqa: F401
                   NORM_L2, SfmmError)
from .features import KEYPOINT_DTYPE, OrbExtractor, extract_features  # noqa: F401
from .matcher import GroupMatcher, Matcher  # noqa: F401

__all__ = ["Matcher", "GroupMatcher", "OrbExtractor", "extract_features", "KEYPOINT_DTYPE", "SfmmError", "DMATCH_DTYPE", "NORM_HAMMING", "NORM_L2", "FLOAT_AUTO", "FLOAT_EXACT",
           "FLOAT_TENSOR", "BINARY_AUTO", "BINARY_POPC", "BINARY_TENSOR"]
