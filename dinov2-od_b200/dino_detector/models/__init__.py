"""Module containers of the libdod detector.

The three public names are the ones callers of the reference import from `dino_detector.models`
(train.py:20, dino_detector/__init__.py:2); they are resolved lazily so that importing the package
does not pull in torch-heavy submodules until a class is actually used.
"""
import importlib

_EXPORTS = {
    "DETRDecoder": "detr_decoder",
    "DINOv2Backbone": "dinov2_backbone",
    "DINOv2ObjectDetector": "detector",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    try:
        submodule = _EXPORTS[name]
    except KeyError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
    value = getattr(importlib.import_module(f"{__name__}.{submodule}"), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + __all__)
