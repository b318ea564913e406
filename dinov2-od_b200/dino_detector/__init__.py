"""B200-native drop-in for mudit1729/dinov2-od's `dino_detector` hot path.

Same import surface as the reference package (reference dino_detector/__init__.py:2):
    from dino_detector.models import DINOv2ObjectDetector, DINOv2Backbone, DETRDecoder
    from dino_detector.matching import HungarianMatcher, build_matcher
    from dino_detector.losses import SetCriterion, build_criterion
All math runs in libdod.so (hand-written sm_100a kernels behind a C ABI, include/dod.h).

Overlay onto a reference checkout
---------------------------------
Only the hot path lives here (models/, matching.py, losses.py, the model/box helpers and
`evaluate_coco` of utils.py, config.py).  The reference's callers -- `train.py`, `validate.py`,
`dataset.py` and the logging / TensorBoard / pycocotools glue of its `utils.py:243-384` -- stay the
reference's own, unmodified files: when `DOD_REFERENCE_DIR` names a reference checkout (the directory
that contains `dino_detector/`), its package directory is appended to this package's `__path__`, so

    DOD_REFERENCE_DIR=/path/to/dinov2-od PYTHONPATH=/path/to/repo/dinov2-od_b200 \
        python -m dino_detector.train --lightweight ...

runs the reference's `train.py` (`from dino_detector.models.detector import DINOv2ObjectDetector`,
train.py:19-27) against the modules of this package; submodules that exist here shadow the
reference's, everything else resolves to the reference file.  `utils.py` forwards the names it does
not define itself to the reference's `utils.py` the same way.
"""
import os as _os


def reference_package_dir():
    """Directory of the reference's `dino_detector` package (DOD_REFERENCE_DIR, else a `baseline/_ref`
    copy next to this repo's bench.py), or None."""
    here = _os.path.dirname(_os.path.abspath(__file__))
    cands = []
    env = _os.environ.get("DOD_REFERENCE_DIR")
    if env:
        cands += [_os.path.join(env, "dino_detector"), env]
    cands.append(_os.path.join(_os.path.dirname(_os.path.dirname(here)), "baseline", "_ref", "dino_detector"))
    for c in cands:
        if _os.path.isfile(_os.path.join(c, "train.py")) and _os.path.abspath(c) != here:
            return _os.path.abspath(c)
    return None


_ref = reference_package_dir()
if _ref is not None and _ref not in __path__:
    __path__.append(_ref)          # after ours: only modules this package lacks come from the reference
del _ref

from .models import DINOv2ObjectDetector  # noqa: E402,F401  (reference dino_detector/__init__.py:2)
