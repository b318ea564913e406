"""CPU oracle for the COCO post-processing -- TEST INFRASTRUCTURE ONLY.
Restates the python loops of reference dino_detector/utils.py:195-233 (torch CPU + python floats)."""
import torch


def coco_detections(pred_logits, pred_boxes, image_ids, threshold=0.05):
    results = []
    scores = torch.sigmoid(pred_logits.float().cpu())
    boxes = pred_boxes.float().cpu()
    for i in range(scores.shape[0]):
        img_scores, b = scores[i], boxes[i]
        xyxy = torch.stack([b[:, 0] - 0.5 * b[:, 2], b[:, 1] - 0.5 * b[:, 3],
                            b[:, 0] + 0.5 * b[:, 2], b[:, 1] + 0.5 * b[:, 3]], dim=-1)       # utils.py:73-92
        for cls_idx in range(img_scores.shape[1]):
            if cls_idx == 0:
                continue
            cls_scores = img_scores[:, cls_idx]
            keep = cls_scores > threshold
            if not keep.any():
                continue
            for score, box in zip(cls_scores[keep].numpy(), xyxy[keep].numpy()):
                x1, y1, x2, y2 = box
                results.append({"image_id": int(image_ids[i]), "category_id": int(cls_idx),
                                "bbox": [float(x1), float(y1), float(x2 - x1), float(y2 - y1)],
                                "score": float(score)})
    return results
