"""GPU parity of HungarianMatcher: indices vs the reference's own outputs (tests/golden), the
CPU oracle, and scipy on our fp32 cost matrix; BASELINE config 3 shapes."""
import numpy as np
import pytest
import torch

from helpers import golden, manifest, matcher_oracle, synth

pytestmark = pytest.mark.gpu


def _cuda_targets(targets):
    return [{k: v.cuda() for k, v in t.items()} for t in targets]


@pytest.mark.parametrize("tag", ["q100", "q25", "dupes"])
def test_matcher_matches_reference_golden(tag):
    from dino_detector.matching import HungarianMatcher
    man = manifest()["matcher_" + tag]
    preds = synth.make_predictions(man["batch"], man["queries"], seed=man["pred_seed"])
    targets = synth.make_targets(man["batch"], max_gt=man["max_gt"], seed=man["target_seed"])
    if tag == "dupes":
        for t in targets:
            if len(t["labels"]) >= 2:
                t["boxes"][1] = t["boxes"][0]
                t["labels"][1] = t["labels"][0]
    m = HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    idx = m({k: v.cuda() for k, v in preds.items()}, _cuda_targets(targets))
    g = golden("matcher_" + tag)
    assert len(idx) == man["batch"]
    for i, (a, b) in enumerate(idx):
        assert a.dtype == torch.int64 and b.dtype == torch.int64 and not a.is_cuda
        assert np.array_equal(a.numpy(), g[f"i{i}"]), (tag, i)
        assert np.array_equal(b.numpy(), g[f"j{i}"]), (tag, i)


@pytest.mark.parametrize("compat", [True, False])
def test_config3_bit_exact_vs_scipy(compat):
    """BASELINE config 3: batch 256, 100 queries x up to 50 GT, class + L1 + GIoU cost."""
    from scipy.optimize import linear_sum_assignment
    from dino_detector.matching import HungarianMatcher
    preds = synth.make_predictions(256, 100, seed=0)
    targets = synth.make_targets(256, max_gt=50, seed=0)
    m = HungarianMatcher()
    m.reference_compat = compat
    dev_preds = {k: v.cuda() for k, v in preds.items()}
    out_q, out_t, status, counts, cost = m.match_device(dev_preds, targets)   # CPU targets are accepted too
    idx = m(dev_preds, targets)
    cost_h = cost.cpu().numpy()
    want = matcher_oracle.match(preds["pred_logits"], preds["pred_boxes"], targets, reference_compat=compat)
    n_same_as_oracle = 0
    for b, t in enumerate(targets):
        n = len(t["labels"])
        ri, ci = linear_sum_assignment(cost_h[b, :, :n])
        assert np.array_equal(idx[b][0].numpy(), ri) and np.array_equal(idx[b][1].numpy(), ci), b
        n_same_as_oracle += int(np.array_equal(want[b][0].numpy(), ri) and np.array_equal(want[b][1].numpy(), ci))
        c_ref = matcher_oracle.cost_matrix(preds["pred_logits"][0 if compat else b], preds["pred_boxes"][0 if compat else b],
                                           t["labels"], t["boxes"]).numpy()
        assert n == 0 or np.abs(cost_h[b, :, :n] - c_ref).max() < 2e-5
    # the CPU-computed cost differs from ours by ulps; assignments still agree (ties aside)
    assert n_same_as_oracle >= 254


def test_empty_targets_and_all_empty():
    from dino_detector.matching import HungarianMatcher
    preds = {k: v.cuda() for k, v in synth.make_predictions(3, 10, seed=1).items()}
    targets = synth.make_targets(3, max_gt=4, seed=7, min_gt=1)
    targets[1] = {"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros((0, 4))}
    idx = HungarianMatcher()(preds, targets)
    assert idx[1][0].numel() == 0 and idx[1][1].numel() == 0
    assert idx[0][0].numel() == len(targets[0]["labels"])
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros((0, 4))}] * 3
    idx = HungarianMatcher()(preds, empty)
    assert all(i.numel() == 0 and j.numel() == 0 for i, j in idx)


def test_nan_cost_raises_like_scipy():
    from dino_detector.matching import HungarianMatcher
    preds = {k: v.cuda() for k, v in synth.make_predictions(2, 10, seed=1).items()}
    preds["pred_boxes"][0, 3, 2] = float("nan")
    targets = synth.make_targets(2, max_gt=4, seed=7, min_gt=2)
    with pytest.raises(ValueError):
        HungarianMatcher()(preds, targets)


def test_zero_weights_assert():
    from dino_detector.matching import HungarianMatcher
    with pytest.raises(AssertionError):
        HungarianMatcher(cost_class=0, cost_bbox=0, cost_giou=0)
