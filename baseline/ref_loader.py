"""Load the UNMODIFIED reference (mudit1729/dinov2-od) from baseline/_ref for the bench's reference arm.

baseline/_ref is created once, offline, by

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference>

(`install()` below; `__graft_entry__.build()` runs it when /root/reference is present).  The directory is
git-ignored but travels to the GPU box with the gpurun snapshot.  Nothing under baseline/_ref is edited;
the two shims of SURVEY.md 8(c) are applied from the OUTSIDE:

  1. `pycocotools{,.coco,.cocoeval}` stub modules if the real package is missing (the reference's
     utils.py:5-6 imports them at module top; they are only used by the offline COCO metric code);
  2. `transformers.Dinov2Model.from_pretrained` builds the named architecture from a `Dinov2Config`
     (image_size 518, patch 14) with HF's random init instead of downloading a checkpoint (no network).

The reference's modules use relative imports only, so the package is imported under the alias
`dino_detector_reference` and can live in one process next to this repo's `dino_detector`.

MEASUREMENT / TEST INFRASTRUCTURE ONLY: imported by bench.py (reference arm, cpu_baseline,
torch_eager_b200 record) and tests/; never by the product package.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
REF_PKG = os.path.join(REF_ROOT, "dino_detector")
ALIAS = "dino_detector_reference"

_HF_VARIANTS = {
    "small": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6),
    "base": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12),
    "large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16),
    "giant": dict(hidden_size=1536, num_hidden_layers=40, num_attention_heads=24, use_swiglu_ffn=True),
}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "models", "detector.py"))


def install(source="/root/reference", force=False) -> bool:
    """pip-install the reference into baseline/_ref from a writable copy (the checkout is read-only)."""
    if available() and not force:
        return True
    if not os.path.isdir(source):
        return False
    tmp = tempfile.mkdtemp(prefix="dod_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(source, src)
        for dp, _, fs in os.walk(src):
            os.chmod(dp, 0o755)
            for f in fs:
                os.chmod(os.path.join(dp, f), 0o644)
        if force and os.path.isdir(REF_ROOT):
            shutil.rmtree(REF_ROOT)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_ROOT, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference install failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


def stub_missing_modules(extra=()):
    """Shim 1: stub modules for imports the reference makes at module top but never uses on the hot path."""
    made = []
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", *extra):
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError):
            pass
        sys.modules[name] = types.ModuleType(name)
        made.append(name)
    if "pycocotools.coco" in made:
        sys.modules["pycocotools.coco"].COCO = object
    if "pycocotools.cocoeval" in made:
        sys.modules["pycocotools.cocoeval"].COCOeval = object
    return made


def patch_from_pretrained(layer_override=None):
    """Shim 2: Dinov2Model.from_pretrained -> random-init model of the named size (no hub access)."""
    from transformers import Dinov2Config, Dinov2Model

    def from_config(name, *a, **k):
        variant = next((v for v in _HF_VARIANTS if v in name), "base")
        cfg = dict(_HF_VARIANTS[variant])
        if layer_override:
            cfg["num_hidden_layers"] = layer_override
        return Dinov2Model(Dinov2Config(image_size=518, patch_size=14, **cfg))

    Dinov2Model.from_pretrained = staticmethod(from_config)


def load(alias=ALIAS):
    """-> the reference package imported under `alias` (models, matching, losses, utils reachable as
    attributes after import_module)."""
    if not available():
        raise ImportError("baseline/_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` in the "
                          "build container (needs /root/reference)")
    if alias in sys.modules:
        return sys.modules[alias]
    stub_missing_modules()
    patch_from_pretrained()
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF_PKG, "__init__.py"),
                                                  submodule_search_locations=[REF_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        for k in [k for k in sys.modules if k == alias or k.startswith(alias + ".")]:
            sys.modules.pop(k, None)
        raise
    for sub in ("matching", "losses", "utils"):
        importlib.import_module(f"{alias}.{sub}")
    return mod


def build_detector(quiet=True, **ctor):
    """The reference's DINOv2ObjectDetector(**ctor) with its own constructor prints silenced."""
    import contextlib
    import io
    ref = load()
    with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
        return ref.models.DINOv2ObjectDetector(**ctor)
