"""GPU parity of the fused SetCriterion: losses and gradients vs the reference's own outputs
(tests/golden/criterion_*.npz) and the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import criterion_oracle, golden, manifest, matcher_oracle, synth

pytestmark = pytest.mark.gpu
WEIGHTS = {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}


def _build():
    from dino_detector.losses import SetCriterion
    from dino_detector.matching import HungarianMatcher
    return SetCriterion(HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2), 91, dict(WEIGHTS))


@pytest.mark.parametrize("tag", ["q100", "q25"])
def test_losses_and_gradients_match_reference_golden(tag):
    man = manifest()["criterion_" + tag]
    preds = synth.make_predictions(man["batch"], man["queries"], seed=man["pred_seed"])
    preds = {k: v.cuda().requires_grad_(True) for k, v in preds.items()}
    targets = synth.make_targets(man["batch"], max_gt=man["max_gt"], seed=man["target_seed"])
    crit = _build()
    ld = crit(preds, [{k: v.cuda() for k, v in t.items()} for t in targets])
    assert set(ld) == {"loss_ce", "loss_bbox", "loss_giou"}
    sum(ld.values()).backward()
    g = golden("criterion_" + tag)
    for k in ld:
        assert abs(ld[k].item() - float(g[k])) < 2e-5 * max(1.0, abs(float(g[k]))), k
    assert np.abs(preds["pred_logits"].grad.cpu().numpy() - g["dlogits"]).max() < 1e-5 * np.abs(g["dlogits"]).max() + 1e-7
    assert np.abs(preds["pred_boxes"].grad.cpu().numpy() - g["dboxes"]).max() < 1e-4 * np.abs(g["dboxes"]).max() + 1e-7


def test_separate_loss_scales_and_per_image_mode():
    """Upstream scalars differ per loss (weighted sum) and reference_compat=False (per-image rows)."""
    preds = synth.make_predictions(5, 50, seed=9)
    targets = synth.make_targets(5, max_gt=20, seed=10)
    crit = _build()
    crit.matcher.reference_compat = False
    p_gpu = {k: v.cuda().requires_grad_(True) for k, v in preds.items()}
    ld = crit(p_gpu, targets)
    (0.5 * ld["loss_ce"] + 2.0 * ld["loss_bbox"] + 3.0 * ld["loss_giou"]).backward()
    p_cpu = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    idx = matcher_oracle.match(p_cpu["pred_logits"].detach(), p_cpu["pred_boxes"].detach(), targets, reference_compat=False)
    ref = criterion_oracle.set_criterion(p_cpu["pred_logits"], p_cpu["pred_boxes"], targets, idx, num_classes=91)
    (0.5 * ref["loss_ce"] + 2.0 * ref["loss_bbox"] + 3.0 * ref["loss_giou"]).backward()
    for k in ref:
        assert abs(ld[k].item() - ref[k].item()) < 2e-5 * max(1.0, abs(ref[k].item()))
    assert (p_gpu["pred_logits"].grad.cpu() - p_cpu["pred_logits"].grad).abs().max() < 1e-5 * p_cpu["pred_logits"].grad.abs().max()
    assert (p_gpu["pred_boxes"].grad.cpu() - p_cpu["pred_boxes"].grad).abs().max() < 1e-4 * p_cpu["pred_boxes"].grad.abs().max()


def test_no_targets_gives_pure_background_loss():
    preds = {k: v.cuda().requires_grad_(True) for k, v in synth.make_predictions(2, 10, seed=1).items()}
    empty = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros((0, 4))}] * 2
    ld = _build()(preds, empty)
    assert ld["loss_bbox"].item() == 0.0 and ld["loss_giou"].item() == 0.0 and ld["loss_ce"].item() > 0
    sum(ld.values()).backward()
    assert preds["pred_boxes"].grad.abs().max().item() == 0.0


def test_full_train_step_detector_plus_criterion():
    """model -> criterion -> backward -> Adam, the loop of reference train.py:1075-1110."""
    from helpers import build_product_model
    model, sd, kw = build_product_model("c1_small_deform", device="cuda", dropout=0.1)
    model.train()
    crit = _build()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=2e-4, weight_decay=1e-4)
    x = synth.make_images(2, 224, 224, seed=3).cuda()
    targets = synth.make_targets(2, max_gt=8, seed=12, min_gt=2)
    first = last = None
    for it in range(6):
        opt.zero_grad()
        ld = crit(model(x), targets)
        loss = sum(ld.values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    assert np.isfinite(last) and last < first
