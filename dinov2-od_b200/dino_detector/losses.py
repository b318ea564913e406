"""SetCriterion on the GPU (reference dino_detector/losses.py:71-254).

Same constructor and `forward(outputs, targets) -> {"loss_ce", "loss_bbox", "loss_giou"}` contract
(weights from `weight_dict` applied, like the reference).  The matcher's device-resident assignment
feeds ONE fused kernel (dod_criterion) that evaluates the focal classification loss, the L1 and GIoU
box losses AND their gradients, so the train step has no per-image python loop, no one-hot tensor and
-- with `strict=False` -- no device->host sync at all.  `num_boxes` follows the reference quirk:
all-reduced with SUM and NOT divided by the world size (losses.py:228-230).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .matching import HungarianMatcher


class _CriterionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, boxes, crit, packed, out_q, out_t, num_boxes):
        labels, tboxes, offsets, _, _ = packed
        wd = crit.weight_dict
        losses, dlogits, dboxes, dboxes_giou = ops.criterion(
            logits.detach().float().contiguous(), boxes.detach().float().contiguous(), labels, tboxes, offsets,
            out_q, out_t, num_boxes, alpha=float(crit.focal_alpha), gamma=float(crit.focal_gamma),
            w_ce=float(wd.get("loss_ce", 1.0)), w_bbox=float(wd.get("loss_bbox", 1.0)),
            w_giou=float(wd.get("loss_giou", 1.0)))
        ctx.save_for_backward(dlogits, dboxes, dboxes_giou)
        ctx.in_dtypes = (logits.dtype, boxes.dtype)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_ce, g_bbox, g_giou):
        dlogits, dboxes, dboxes_giou = ctx.saved_tensors
        b, q, c = dlogits.shape
        dev = dlogits.device
        zero = torch.zeros((), device=dev)
        g = torch.stack([x if x is not None else zero for x in (g_ce, g_bbox, g_giou)]).float()
        gl = ops.eltwise(ops.ELT_AXPBY, dlogits.view(b * q, c), vec=g[0:1].contiguous(), out_dtype=torch.float32)
        gb = ops.eltwise(ops.ELT_AXPBY, dboxes.view(b * q, 4), dboxes_giou.view(b * q, 4),
                         vec=g[1:3].contiguous(), out_dtype=torch.float32)
        return gl.view(b, q, c).to(ctx.in_dtypes[0]), gb.view(b, q, 4).to(ctx.in_dtypes[1]), None, None, None, None, None


class SetCriterion(nn.Module):
    def __init__(self, matcher, num_classes, weight_dict, focal_alpha=0.25, focal_gamma=2.0):
        super().__init__()
        self.matcher = matcher
        self.num_classes = num_classes
        self.weight_dict = weight_dict
        self.focal_alpha = focal_alpha
        self.focal_gamma = focal_gamma
        # strict: check the solver status on the host (one tiny D2H sync) and raise ValueError like
        # scipy does for NaN / infeasible costs; False keeps the whole step free of host syncs.
        self.strict = True
        # losses.py:228-230 all-reduces num_boxes whenever torch.distributed is initialised; False keeps this
        # rank's count (a single-process evaluation of a whole batch inside a multi-rank job, e.g. a parity check)
        self.sync_num_boxes = True

    def forward(self, outputs, targets):
        dev = outputs["pred_logits"].device
        if not isinstance(self.matcher, HungarianMatcher):
            raise TypeError("SetCriterion needs the libdod HungarianMatcher (device-resident assignment)")
        packed = self.matcher.pack_targets(targets, dev)
        # losses.py:225-230: num_boxes = sum of GT counts
        num_boxes = torch.tensor([float(sum(packed[3]))], dtype=torch.float32).to(dev, non_blocking=True)
        return self.forward_packed(outputs, packed, num_boxes)

    def forward_packed(self, outputs, packed, num_boxes):
        """Device-only form (no host reads; capturable in a CUDA graph when strict is False):
        packed = (labels i64 [T], boxes f32 [T, 4], offsets i32 [B + 1], per-image counts or None,
        max targets per image); num_boxes = f32 [1] device tensor with this rank's GT count."""
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        out_q, out_t, status, _, _ = self.matcher.match_device(outputs, None, packed=packed)
        if self.strict and bool((status != 0).any()):
            raise ValueError("matrix contains invalid numeric entries")
        # losses.py:228-230: all-reduced with SUM (no / world), clamp(min=1)
        if self.sync_num_boxes and dist.is_available() and dist.is_initialized():
            num_boxes = num_boxes.clone()
            dist.all_reduce(num_boxes)
        num_boxes = torch.clamp(num_boxes, min=1)
        l_ce, l_bbox, l_giou = _CriterionFn.apply(logits, boxes, self, packed, out_q, out_t, num_boxes)
        return {"loss_ce": l_ce, "loss_bbox": l_bbox, "loss_giou": l_giou}


def build_criterion(matcher, num_classes, weight_dict, focal_alpha=0.25, focal_gamma=2.0):
    """reference losses.py:244-254."""
    return SetCriterion(matcher=matcher, num_classes=num_classes, weight_dict=weight_dict,
                        focal_alpha=focal_alpha, focal_gamma=focal_gamma)
