"""Model / box utilities with the reference's names (reference dino_detector/utils.py:14-164).

`MLP`, `LoraLinear` and `add_lora_to_module` keep the reference's parameter names so that
state_dict keys match; they are parameter containers -- the detector's forward reads
their tensors and runs libdod kernels, it does not call these modules' `forward`.
The box helpers are thin torch expressions kept for API compatibility (the matcher and
criterion hot paths use the fused kernels instead).

Names of the reference's utils.py that are NOT on the hot path -- `compute_coco_metrics`,
`setup_logger`, `setup_tensorboard`, `log_metrics`, `log_images` (reference utils.py:243-384, imported by
train.py:22-25) -- are not re-implemented: module `__getattr__` forwards them to the reference's own
`utils.py` from the overlaid checkout (DOD_REFERENCE_DIR, see the package docstring).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class MLP(nn.Module):
    """reference utils.py:14-30 -- keys `mlp.0.*`, `mlp.2.*`."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        layers = []
        for i in range(num_layers):
            in_dim = input_dim if i == 0 else hidden_dim
            out_dim = output_dim if i == num_layers - 1 else hidden_dim
            layers.append(nn.Linear(in_dim, out_dim))
            if i < num_layers - 1:
                layers.append(nn.ReLU())
        self.mlp = nn.Sequential(*layers)


class LoraLinear(nn.Module):
    """reference utils.py:46-70: y = W x + b + alpha * B(A x); alpha is not divided by r.
    Keys: `linear.{weight,bias}` (frozen), `lora_A.weight [r,in]`, `lora_B.weight [out,r]`."""

    def __init__(self, linear_layer: nn.Linear, r=4, alpha=1.0):
        super().__init__()
        self.linear = linear_layer
        self.in_features = linear_layer.in_features
        self.out_features = linear_layer.out_features
        self.r = r
        self.alpha = alpha
        self.lora_A = nn.Linear(self.in_features, r, bias=False)
        self.lora_B = nn.Linear(r, self.out_features, bias=False)
        nn.init.zeros_(self.lora_B.weight)
        for param in self.linear.parameters():
            param.requires_grad = False


def add_lora_to_module(module, r=4, alpha=1.0):
    """reference utils.py:33-43: recursively wrap every nn.Linear child in a LoraLinear."""
    for name, child in module.named_children():
        add_lora_to_module(child, r=r, alpha=alpha)
        if isinstance(child, nn.Linear):
            setattr(module, name, LoraLinear(child, r=r, alpha=alpha))


def box_cxcywh_to_xyxy(x):
    """reference utils.py:73-92."""
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([x_c - 0.5 * w, y_c - 0.5 * h, x_c + 0.5 * w, y_c + 0.5 * h], dim=-1)


def box_xyxy_to_cxcywh(x):
    """reference utils.py:95-108."""
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


def box_area(boxes):
    """reference utils.py:111-121."""
    return (boxes[..., 2] - boxes[..., 0]) * (boxes[..., 3] - boxes[..., 1])


def generalized_box_iou(boxes1, boxes2):
    """reference utils.py:124-164 (pairwise, no eps)."""
    area1, area2 = box_area(boxes1), box_area(boxes2)
    lt = torch.max(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    iou = inter / union
    lt_e = torch.min(boxes1[:, None, :2], boxes2[:, :2])
    rb_e = torch.max(boxes1[:, None, 2:], boxes2[:, 2:])
    wh_e = (rb_e - lt_e).clamp(min=0)
    area_e = wh_e[:, :, 0] * wh_e[:, :, 1]
    return iou - (area_e - union) / area_e


def coco_detections(outputs, image_ids, threshold=0.05):
    """Detector outputs -> list of COCO result dicts, same content and order as the loops at
    reference utils.py:195-233 (class 0 skipped, score > threshold, bbox = [x, y, w, h])."""
    from . import ops
    scores, boxes, classes, counts = ops.postprocess(outputs["pred_logits"], outputs["pred_boxes"], threshold)
    counts = counts.cpu().tolist()                      # one small D2H copy
    results = []
    for b, n in enumerate(counts):
        if n == 0:
            continue
        s, bx, cl = scores[b, :n].cpu().tolist(), boxes[b, :n].cpu().tolist(), classes[b, :n].cpu().tolist()
        img_id = int(image_ids[b])
        results.extend({"image_id": img_id, "category_id": int(c), "bbox": [float(v) for v in box], "score": float(sc)}
                       for sc, box, c in zip(s, bx, cl))
    return results


def evaluate_coco(model, dataloader, device, output_file=None):
    """reference utils.py:167-240 (tqdm progress bar omitted): forward + COCO-format detections."""
    import json
    model.eval()
    results = []
    with torch.no_grad():
        for images, targets in dataloader:
            outputs = model(images.to(device))
            ids = [int(t.get("image_id", i)) for i, t in enumerate(targets)]
            results.extend(coco_detections(outputs, ids))
    if output_file is not None:
        with open(output_file, "w") as f:
            json.dump(results, f)
    return results


# ---------------------------------------------------------------------------
# everything else: the reference's own utils.py (logging / TensorBoard / pycocotools glue)
# ---------------------------------------------------------------------------
_REFERENCE_UTILS = None


def _reference_utils():
    """The reference checkout's utils.py, executed once as `dino_detector._reference_utils`."""
    global _REFERENCE_UTILS
    if _REFERENCE_UTILS is None:
        import importlib.util
        import os
        import sys
        from . import reference_package_dir
        ref = reference_package_dir()
        if ref is None:
            raise ImportError("this name lives in the reference's dino_detector/utils.py: set DOD_REFERENCE_DIR to a "
                              "checkout of mudit1729/dinov2-od (the directory that contains dino_detector/)")
        name = __package__ + "._reference_utils"
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref, "utils.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            sys.modules.pop(name, None)
            raise
        _REFERENCE_UTILS = mod
    return _REFERENCE_UTILS


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    try:
        return getattr(_reference_utils(), name)
    except ImportError as e:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r} ({e})") from None
