"""Shared test helpers: import paths for oracle/ (test infrastructure) and the product package."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dinov2-od_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import criterion_oracle  # noqa: E402
import detector_oracle  # noqa: E402
import matcher_oracle  # noqa: E402
import synth  # noqa: E402


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def build_product_model(case, device="cpu", quiet=True, **override):
    """Our DINOv2ObjectDetector for a synth.CASES entry with the synthetic weights loaded."""
    from dino_detector.models import DINOv2ObjectDetector
    from dino_detector.models import dinov2_backbone as bb
    kw = synth.case_ctor(case)
    kw.update(override)
    bb._LAYER_OVERRIDE = synth.CASES[case].get("backbone_layers")
    try:
        with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            model = DINOv2ObjectDetector(**kw)
    finally:
        bb._LAYER_OVERRIDE = None
    sd = synth.case_state_dict(case)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), sd, kw


def oracle_forward(sd, x, kw):
    return detector_oracle.detector_forward(sd, x, dino_model_name=kw["dino_model_name"], nheads=kw["nheads"],
                                            n_points=kw["n_points"], use_deformable=kw["use_deformable"],
                                            lora_alpha=kw["lora_alpha"])


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def record_margin(case, what, value, tol, **extra):
    """Append the ACHIEVED error of a parity check next to its tolerance to a JSON-lines file
    ($DOD_MARGINS_FILE, default gpurun_out/parity_margins.jsonl; tools/collect_margins.py turns it into
    profiles/rNN_parity_margins.json) -- so that the distance to the bar is on record, not just pass / fail."""
    path = os.environ.get("DOD_MARGINS_FILE", os.path.join(ROOT, "gpurun_out", "parity_margins.jsonl"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "a") as fh:
            fh.write(json.dumps(dict(case=case, what=what, value=float(value), tol=float(tol), **extra)) + "\n")
    except OSError:
        pass


def rel_err_rms(a, b):
    """RMS of the difference over the RMS of the reference (a tighter reading of "relative" than rel_err)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()
