"""Synthetic inputs for the developer tools (no dependency on oracle/, which is test infrastructure)."""
import torch


def make_targets(batch, *, max_gt=50, num_classes=91, seed=0, min_gt=0):
    """COCO-style targets (dataset.py:102-111): labels int64 [n], boxes cxcywh fp32 [n, 4], non-degenerate."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(min_gt, max_gt + 1, (1,), generator=g))
        cxcy = torch.rand((n, 2), generator=g) * 0.6 + 0.2
        wh = torch.rand((n, 2), generator=g) * 0.3 + 0.02
        out.append({"labels": torch.randint(0, num_classes, (n,), generator=g), "boxes": torch.cat([cxcy, wh], dim=1)})
    return out


def make_predictions(batch, num_queries, num_classes=91, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"pred_logits": torch.randn((batch, num_queries, num_classes), generator=g),
            "pred_boxes": torch.rand((batch, num_queries, 4), generator=g) * 0.5 + 0.25}
