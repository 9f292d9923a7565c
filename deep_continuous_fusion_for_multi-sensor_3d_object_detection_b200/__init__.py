"""B200-native continuous-fusion hot path (BEV KNN -> projection -> bilinear gather -> MLP -> K-sum-pool ->
BEV add) and rotated-box post-process, as a drop-in for the TODO at model.py:199-203 and for
Test.get_bboxes / NMS_SAT / NMS_IOU (test.py:88-175) of
Chanuk-Yang/Deep_Continuous_Fusion_for_Multi-Sensor_3D_Object_Detection.

Host side: Python / PyTorch (device memory, streams).  Device side: libcf_b200.so, hand-written sm_100a
kernels behind the C ABI in include/cf_b200.h.  There is no CPU or PyTorch fallback.
"""
from . import dist_util, geometry, synthetic  # noqa: F401  (importable without the CUDA library)
from ._lib import SO_PATH, build, load  # noqa: F401
from .fusion import ContinuousFusion, FrameContext, FusionRunner, fuse_scales, prepare_frames  # noqa: F401
from .postprocess import PostProcess  # noqa: F401
from .model import ObjectDetection_DCF  # noqa: F401
from .loss import LossTotal  # noqa: F401

__all__ = ["ContinuousFusion", "FrameContext", "FusionRunner", "fuse_scales", "prepare_frames", "PostProcess", "ObjectDetection_DCF", "LossTotal", "geometry",
           "synthetic", "build", "load", "SO_PATH"]
