"""B200-native drop-in for mudit1729/dinov2-od's `dino_detector` hot path.

Same import surface as the reference package (reference dino_detector/__init__.py:2):
    from dino_detector.models import DINOv2ObjectDetector, DINOv2Backbone, DETRDecoder
    from dino_detector.matching import HungarianMatcher, build_matcher
    from dino_detector.losses import SetCriterion, build_criterion
All math runs in libdod.so (hand-written sm_100a kernels behind a C ABI, include/dod.h).
"""
