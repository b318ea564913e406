"""Generate tests/golden/*.npz by running the REFERENCE itself (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference (mudit1729/dinov2-od, pure Python) is imported from /root/reference with the
two shims of SURVEY.md 8c: stub `pycocotools` (imported at module top by utils.py:5-6) and a
`Dinov2Model.from_pretrained` that builds the named architecture from a Dinov2Config instead
of downloading a checkpoint.  Synthetic weights from oracle/synth.py are loaded with
strict=True (this pins the state_dict key/shape contract), the reference forward / matcher
run on CPU fp32, and only the small outputs are stored.  The fixtures travel to the GPU
box; this script and /root/reference do not need to.

TEST INFRASTRUCTURE ONLY -- never imported by the product path.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("DOD_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)

import synth  # noqa: E402
import detector_oracle  # noqa: E402
import matcher_oracle  # noqa: E402
import criterion_oracle  # noqa: E402

_LAYER_OVERRIDE = {"n": None}


def import_reference():
    for n in ["pycocotools", "pycocotools.coco", "pycocotools.cocoeval"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    sys.path.insert(0, REFERENCE)
    from transformers import Dinov2Config, Dinov2Model

    def fake_from_pretrained(name, *a, **k):
        v = detector_oracle.variant_of(name)
        c = detector_oracle.VARIANTS[v]
        return Dinov2Model(Dinov2Config(image_size=518, patch_size=14, hidden_size=c["dim"],
                                        num_hidden_layers=_LAYER_OVERRIDE["n"] or c["layers"],
                                        num_attention_heads=c["heads"], use_swiglu_ffn=c["swiglu"]))

    Dinov2Model.from_pretrained = staticmethod(fake_from_pretrained)
    import dino_detector.models as ref_models
    import dino_detector.matching as ref_matching
    import dino_detector.losses as ref_losses
    return ref_models, ref_matching, ref_losses


def main():
    """`python oracle/make_golden.py [case ...]`: no arguments regenerates everything; case names regenerate only
    those detector fixtures (the manifest entries of the others are kept)."""
    os.makedirs(GOLDEN, exist_ok=True)
    ref_models, ref_matching, ref_losses = import_reference()
    torch.set_grad_enabled(False)
    only = set(sys.argv[1:])
    manifest = {}
    man_path = os.path.join(GOLDEN, "manifest.json")
    if only and os.path.exists(man_path):
        with open(man_path) as fh:
            manifest = json.load(fh)
    for name, case in synth.CASES.items():
        if only and name not in only:
            continue
        kw = synth.case_ctor(name)
        _LAYER_OVERRIDE["n"] = case.get("backbone_layers")
        model = ref_models.DINOv2ObjectDetector(**kw).eval()
        sd = synth.case_state_dict(name)
        model.load_state_dict(sd, strict=True)          # pins key names and shapes
        b, (h, w) = case["batch"], case["hw"]
        x = synth.make_images(b, h, w, seed=1)
        out = model(x)
        mem = model.backbone(x)
        ours = detector_oracle.detector_forward(sd, x, dino_model_name=kw["dino_model_name"],
                                                nheads=kw["nheads"], n_points=kw["n_points"],
                                                use_deformable=kw["use_deformable"],
                                                lora_alpha=kw["lora_alpha"])
        dl = (ours["pred_logits"] - out["pred_logits"]).abs().max().item()
        db = (ours["pred_boxes"] - out["pred_boxes"]).abs().max().item()
        print(f"{name}: logits {tuple(out['pred_logits'].shape)} |oracle-ref| logits {dl:.2e} boxes {db:.2e}")
        np.savez_compressed(os.path.join(GOLDEN, f"detector_{name}.npz"),
                            pred_logits=out["pred_logits"].numpy(), pred_boxes=out["pred_boxes"].numpy(),
                            memory_mean=mem.mean(dim=-1).numpy(), memory_cls=mem[:, 0].numpy())
        manifest[name] = dict(ctor={k: v for k, v in kw.items()}, batch=b, hw=[h, w],
                              backbone_layers=case.get("backbone_layers"), image_seed=1, weight_seed=0,
                              n_keys=len(sd))
        del model, sd

    if only:
        with open(man_path, "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        return
    # ---- matcher: reference HungarianMatcher (matching.py:42-122, scipy LSA) ----
    matcher = ref_matching.HungarianMatcher(cost_class=1, cost_bbox=5, cost_giou=2)
    for tag, (bs, q, max_gt) in {"q100": (16, 100, 50), "q25": (8, 25, 50), "dupes": (4, 50, 20)}.items():
        preds = synth.make_predictions(bs, q, seed=3)
        targets = synth.make_targets(bs, max_gt=max_gt, seed=4)
        if tag == "dupes":                              # duplicated GT boxes -> tied columns
            for t in targets:
                if len(t["labels"]) >= 2:
                    t["boxes"][1] = t["boxes"][0]
                    t["labels"][1] = t["labels"][0]
        idx = matcher(preds, targets)
        ours = matcher_oracle.match(preds["pred_logits"], preds["pred_boxes"], targets, reference_compat=True)
        same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(idx, ours))
        print(f"matcher_{tag}: oracle == reference: {same}")
        arrs = {}
        for i, (a, b) in enumerate(idx):
            arrs[f"i{i}"] = a.numpy()
            arrs[f"j{i}"] = b.numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"matcher_{tag}.npz"), **arrs)
        manifest[f"matcher_{tag}"] = dict(batch=bs, queries=q, max_gt=max_gt, pred_seed=3, target_seed=4)
    # ---- criterion: reference SetCriterion (losses.py:71-241) incl. autograd gradients ----
    torch.set_grad_enabled(True)
    for tag, (bs, q, max_gt) in {"q100": (8, 100, 30), "q25": (6, 25, 40)}.items():
        preds = synth.make_predictions(bs, q, seed=5)
        preds = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
        targets = synth.make_targets(bs, max_gt=max_gt, seed=6)
        crit = ref_losses.SetCriterion(matcher, 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
        ld = crit(preds, targets)
        sum(ld.values()).backward()
        with torch.no_grad():
            idx = matcher(preds, targets)
        mine = criterion_oracle.set_criterion(preds["pred_logits"].detach(), preds["pred_boxes"].detach(), targets,
                                              idx, num_classes=91)
        print(f"criterion_{tag}:", {k: float(v) for k, v in ld.items()},
              "oracle diff", max(abs(float(ld[k]) - float(mine[k])) for k in ld))
        np.savez_compressed(os.path.join(GOLDEN, f"criterion_{tag}.npz"),
                            loss_ce=ld["loss_ce"].detach().numpy(), loss_bbox=ld["loss_bbox"].detach().numpy(),
                            loss_giou=ld["loss_giou"].detach().numpy(),
                            dlogits=preds["pred_logits"].grad.numpy(), dboxes=preds["pred_boxes"].grad.numpy())
        manifest[f"criterion_{tag}"] = dict(batch=bs, queries=q, max_gt=max_gt, pred_seed=5, target_seed=6)
    torch.set_grad_enabled(False)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
