"""GPU: the reference's own `train.py` trains this package's detector.

`python -m dino_detector.train --lightweight ...` (the reference file, overlaid) runs two epochs of its hot loop
(train.py:1067-1110: forward, `criterion(outputs, targets)` with the GPU matcher, `loss.backward()` through the
hand-written backward, `clip_grad_norm_`, `torch.optim.Adam.step`) on a synthetic COCO folder, validates each epoch
(:189-227 -> `evaluate_coco` of this package + the reference's `compute_coco_metrics`) and saves the reference's
checkpoint formats (:1278-1294)."""
import math

import pytest

from test_dropin_cpu import run_driver

pytestmark = pytest.mark.gpu


def test_reference_training_loop_runs_on_libdod(tmp_path):
    rep, out, ref = run_driver(tmp_path, "train_gpu", timeout=1500)
    assert rep["train_file"].startswith(ref)
    assert len(rep["losses"]) == 6                      # 12 images / batch 4 x 2 epochs
    assert all(math.isfinite(v) and v > 0 for v in rep["losses"]), rep["losses"]
    assert rep["launches"] > 1000                       # the steps ran on libdod kernels
    assert rep["final_keys"] == 306 and rep["ckpt_has_optimizer"] and rep["ckpt_epoch"] == 1
    assert rep["val_predictions"] and rep["val_metrics"]
    assert "Training complete." in out
