"""Hungarian matcher on the GPU (reference dino_detector/matching.py:9-135).

Same constructor, same `forward(outputs, targets) -> List[(LongTensor, LongTensor)]` contract
(CPU int64 tensors, prediction indices ascending).  The per-image python loop, the O(B^2)
cost recompute and the 256 device->host syncs + scipy calls of the reference are replaced by
two kernel launches (dod_match_cost, dod_lsap_jv) and ONE device->host copy of the indices.
Assignments are bit-identical to scipy.optimize.linear_sum_assignment on the fp32 cost matrix.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import config, ops


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class=1, cost_bbox=5, cost_giou=2, focal_alpha=0.25, focal_gamma=2.0):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        self.focal_alpha = focal_alpha
        self.focal_gamma = focal_gamma
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "at least one cost should be non-zero"
        # reference matching.py:102 slices C[:num_queries] -> the rows of IMAGE 0 are matched against
        # every image's targets.  True reproduces that; False matches image b's own predictions.
        self.reference_compat = config.reference_compat

    @staticmethod
    def pack_targets(targets, device):
        """List of target dicts (dataset.py:102-111) -> CSR tensors on `device`:
        (labels i64 [T], boxes f32 [T, 4], offsets i32 [B+1], per-image counts, max count)."""
        ns = [int(t["labels"].shape[0]) for t in targets]
        offs = [0]
        for n in ns:
            offs.append(offs[-1] + n)
        max_t = max(ns) if ns else 0
        offsets = torch.tensor(offs, dtype=torch.int32).to(device, non_blocking=True)
        labels = tboxes = None
        if max_t > 0:
            labels = torch.cat([t["labels"].reshape(-1) for t in targets]).to(device=device, dtype=torch.int64)
            tboxes = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(device=device, dtype=torch.float32)
            tboxes = tboxes.contiguous()
        return labels, tboxes, offsets, ns, max_t

    @torch.no_grad()
    def match_device(self, outputs, targets, packed=None):
        """-> (out_q, out_t, status, counts, cost): device int32 [B, K] index arrays (first
        counts[b] entries valid), the per-image pair counts (python list) and the cost tensor."""
        logits = outputs["pred_logits"].detach()
        boxes = outputs["pred_boxes"].detach()
        dev = logits.device
        bs, nq = logits.shape[:2]
        logits = logits.float().contiguous()
        boxes = boxes.float().contiguous()
        labels, tboxes, offsets, ns, max_t = packed if packed is not None else self.pack_targets(targets, dev)
        assert offsets.numel() == bs + 1, "one target dict per image"
        cost = ops.match_cost(logits, boxes, labels, tboxes, offsets, max_t,
                              w_class=float(self.cost_class), w_bbox=float(self.cost_bbox),
                              w_giou=float(self.cost_giou), alpha=float(self.focal_alpha),
                              gamma=float(self.focal_gamma), use_image0_rows=self.reference_compat)
        out_q, out_t, status = ops.lsap(cost, offsets, max_t)
        counts = [min(nq, n) for n in ns] if ns is not None else None
        return out_q, out_t, status, counts, cost

    @torch.no_grad()
    def forward(self, outputs, targets):
        out_q, out_t, status, counts, _ = self.match_device(outputs, targets)
        # one D2H copy, one widening to int64 for the whole batch (512 per-image conversions cost more host time
        # than the two kernels at batch 256); the returned index tensors are row slices of that array
        packed = torch.cat([out_q, out_t, status.view(-1, 1)], dim=1).cpu().to(torch.int64)
        k = out_q.shape[1]
        if bool((packed[:, -1] != 0).any()):
            # scipy raises ValueError for NaN / -inf entries or an infeasible matrix (matching.py:105)
            raise ValueError("matrix contains invalid numeric entries")
        return [(packed[b, :c], packed[b, k:k + c]) for b, c in enumerate(counts)]


def build_matcher(args):
    """reference matching.py:125-135."""
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                            cost_giou=args.set_cost_giou, focal_alpha=args.focal_alpha,
                            focal_gamma=args.focal_gamma)
