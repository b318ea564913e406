"""CPU: the documented drop-in works with the REFERENCE'S OWN CALLER.

The reference's unmodified `dino_detector/train.py` (overlaid through DOD_REFERENCE_DIR / baseline/_ref, see the package
docstring) is imported against this repo's package and driven through `main()` with `--device cpu --lightweight
--only_evaluate`: argument parsing, model construction through the lightweight branch (train.py:603-654), criterion
creation (:159-187, `criterion.to(device)`), logger / TensorBoard setup (reference utils.py:283-345 reached through
`dino_detector.utils`), the shape-filtered checkpoint load with 'module.' prefix handling (:686-747) and
validate -> evaluate_coco -> compute_coco_metrics (:189-227) on an empty COCO folder (a forward on the CPU must fail:
there is no CPU fallback, tests/test_host_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


def reference_dir():
    for c in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(c, "dino_detector", "train.py")):
            return c
    return None


def run_driver(tmp_path, mode, timeout=900):
    ref = reference_dir()
    if ref is None:
        pytest.skip("no reference checkout (/root/reference or baseline/_ref)")
    env = dict(os.environ, DOD_REFERENCE_DIR=ref)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), str(tmp_path), mode],
                       capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    with open(os.path.join(tmp_path, "report.json")) as fh:
        return json.load(fh), r.stdout + r.stderr, ref


def test_reference_train_py_runs_against_this_package(tmp_path):
    rep, out, ref = run_driver(tmp_path, "evaluate_cpu")
    # the callers are the reference's files, the hot-path classes are ours
    for k in ("train_file", "dataset_file", "validate_file"):
        assert rep[k].startswith(ref), rep[k]
    assert rep["setup_logger_module"] == "dino_detector._reference_utils"
    # checkpoint filter (train.py:712-738): 306 keys, 'module.' prefix stripped, the mis-shaped one dropped
    assert rep["n_keys"] == 306 and rep["keys"] == 305 and not rep["has_bad"] and rep["bias_loaded"]
    assert "Loaded 305 compatible parameters from checkpoint" in out
    assert "Trainable parameters: 841,979" in out      # the reference's logger; count as probed in SURVEY section 11
    assert rep["metrics_written"]


def test_utils_forwards_cold_path_names_to_the_reference(tmp_path):
    ref = reference_dir()
    if ref is None:
        pytest.skip("no reference checkout")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import dropin_driver as d; d.install_stubs();"
            "from dino_detector.utils import (evaluate_coco, compute_coco_metrics, setup_logger, setup_tensorboard,"
            " log_metrics, log_images, MLP, LoraLinear, add_lora_to_module, box_cxcywh_to_xyxy, generalized_box_iou);"
            "import dino_detector.utils as u;"
            "assert evaluate_coco.__module__ == 'dino_detector.utils' and MLP.__module__ == 'dino_detector.utils';"
            "assert log_images.__module__ == 'dino_detector._reference_utils';"
            "import pytest\n"
            "try:\n    u.no_such_name\nexcept AttributeError: print('ok')"
            % (os.path.join(ROOT, "dinov2-od_b200"), os.path.join(ROOT, "tests")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, DOD_REFERENCE_DIR=ref))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
