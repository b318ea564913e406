"""Runs the REFERENCE's own caller (`dino_detector/train.py`, unmodified, from the overlaid checkout) against this
repo's `dino_detector` package in a fresh interpreter.  Used by tests/test_dropin_cpu.py and tests/test_dropin_gpu.py.

    python tests/dropin_driver.py WORKDIR MODE        MODE = evaluate_cpu | train_gpu

The reference imports matplotlib (train.py:39) and pycocotools (its utils.py:5-6) at module top; neither is in this
image, so functional stand-ins are installed first (a user's environment has the real ones).  Everything else --
argument parsing, model construction through the `--lightweight` branch (train.py:603-654), criterion creation
(:159-187), checkpoint filter (:686-747), the hot loop (:1067-1110), validate (:189-227), checkpoint save (:1278-1294)
-- is the reference's code.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_stubs():
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "plot", "xlabel", "ylabel", "title", "legend", "grid", "savefig", "close", "subplot",
                 "tight_layout"):
        setattr(plt, name, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    class COCO:
        def __init__(self, annotation_file=None):
            with open(annotation_file) as fh:
                self.dataset = json.load(fh)

        def loadRes(self, results):
            res = COCO.__new__(COCO)
            res.dataset = {"annotations": list(results)}
            return res

    class COCOeval:
        def __init__(self, gt, dt, iou_type):
            self.stats = [0.0] * 12
            self.n_dt = len(dt.dataset["annotations"])

        def evaluate(self):
            pass

        def accumulate(self):
            pass

        def summarize(self):
            print(f"[stub COCOeval] {self.n_dt} detections")

    pk = types.ModuleType("pycocotools")
    pc, pe = types.ModuleType("pycocotools.coco"), types.ModuleType("pycocotools.cocoeval")
    pc.COCO, pe.COCOeval = COCO, COCOeval
    for n, m in (("pycocotools", pk), ("pycocotools.coco", pc), ("pycocotools.cocoeval", pe)):
        sys.modules.setdefault(n, m)


def make_coco_folder(path, n_images, seed=0):
    """A tiny COCO-format folder: PNG images + annotation json (dataset.py:9-35 reads exactly these fields)."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    os.makedirs(path, exist_ok=True)
    images, anns = [], []
    for i in range(n_images):
        w, h = 96 + 8 * (i % 3), 80
        Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(os.path.join(path, f"{i + 1:06d}.png"))
        images.append({"id": i + 1, "file_name": f"{i + 1:06d}.png", "width": w, "height": h})
        for k in range(1 + i % 3):
            bw, bh = float(rng.uniform(10, 40)), float(rng.uniform(10, 40))
            x, y = float(rng.uniform(0, w - bw)), float(rng.uniform(0, h - bh))
            anns.append({"id": len(anns) + 1, "image_id": i + 1, "category_id": int(rng.integers(1, 5)),
                         "bbox": [x, y, bw, bh], "area": bw * bh, "iscrowd": 0})
    ann_file = os.path.join(path, "annotations.json")
    with open(ann_file, "w") as fh:
        json.dump({"images": images, "annotations": anns,
                   "categories": [{"id": c, "name": f"c{c}"} for c in range(1, 6)]}, fh)
    return ann_file


def main():
    work, mode = sys.argv[1], sys.argv[2]
    install_stubs()
    sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
    import torch
    import dino_detector
    from dino_detector import config
    ref_dir = dino_detector.reference_package_dir()
    assert ref_dir is not None, "no reference checkout to overlay (DOD_REFERENCE_DIR / baseline/_ref)"
    config.num_workers = 0                 # config.py is how a user changes these (train.py:28-37 imports them)
    config.num_epochs = 2
    from dino_detector import train        # the reference's train.py, found through the overlaid __path__
    import dino_detector.dataset as ds
    import dino_detector.validate as val
    ours = os.path.join(ROOT, "dinov2-od_b200", "dino_detector")
    report = {"train_file": train.__file__, "dataset_file": ds.__file__, "validate_file": val.__file__,
              "detector_file": sys.modules[train.DINOv2ObjectDetector.__module__].__file__,
              "criterion_file": sys.modules[train.SetCriterion.__module__].__file__,
              "matcher_file": sys.modules[train.HungarianMatcher.__module__].__file__,
              "setup_logger_module": train.setup_logger.__module__,
              "compute_coco_metrics_module": train.compute_coco_metrics.__module__}
    assert os.path.dirname(train.__file__) == ref_dir, report
    for k in ("detector_file", "criterion_file", "matcher_file"):
        assert report[k].startswith(ours), report
    out_dir = os.path.join(work, "out")
    if mode == "evaluate_cpu":
        # checkpoint in the reference's format (train.py:1281-1287) written from a model of THIS package, with a
        # 'module.' prefix (DDP checkpoint into a non-DDP model, :697-705) and one shape-mismatched tensor (:712-722)
        from dino_detector.models import DINOv2ObjectDetector
        kw = dict(num_classes=91, dino_model_name="facebook/dinov2-small", hidden_dim=256, num_queries=25,
                  num_decoder_layers=2, dim_feedforward=512, lora_r=1, nheads=4)        # train.py:631-641
        torch.manual_seed(0)
        sd = DINOv2ObjectDetector(**kw).state_dict()
        marker = torch.full_like(sd["decoder.class_embed.bias"], 0.125)
        sd["decoder.class_embed.bias"] = marker
        sd["decoder.query_embed.weight"] = torch.zeros(7, 256)                          # wrong shape -> filtered
        ckpt = os.path.join(work, "ckpt.pth")
        torch.save({"epoch": 4, "model_state_dict": {"module." + k: v for k, v in sd.items()},
                    "metrics_history": {"epochs": [], "train_loss": [], "val_epochs": [], "val_ap": [],
                                        "val_ap50": [], "val_ap75": []}}, ckpt)
        val_dir = os.path.join(work, "val")
        ann = make_coco_folder(val_dir, 0)
        sys.argv = ["train.py", "--device", "cpu", "--lightweight", "--only_evaluate", "--val_images", val_dir,
                    "--val_annotations", ann, "--checkpoint", ckpt, "--output_dir", out_dir, "--batch_size", "2"]
        captured = {}
        real_load = torch.nn.Module.load_state_dict

        def spy(self, state_dict, *a, **k):
            if type(self).__name__ == "DINOv2ObjectDetector":
                captured["keys"] = len(state_dict)
                captured["has_bad"] = "decoder.query_embed.weight" in state_dict
            r = real_load(self, state_dict, *a, **k)
            if type(self).__name__ == "DINOv2ObjectDetector":
                captured["bias_loaded"] = bool(torch.equal(self.state_dict()["decoder.class_embed.bias"], marker))
            return r

        torch.nn.Module.load_state_dict = spy
        try:
            train.main()
        finally:
            torch.nn.Module.load_state_dict = real_load
        report.update(captured, n_keys=len(sd),
                      metrics_written=os.path.exists(os.path.join(out_dir, "val_metrics_epoch_0.json")))
    elif mode == "train_gpu":
        tr_dir, va_dir = os.path.join(work, "train"), os.path.join(work, "val")
        tr_ann, va_ann = make_coco_folder(tr_dir, 12, seed=1), make_coco_folder(va_dir, 4, seed=2)
        sys.argv = ["train.py", "--lightweight", "--train_images", tr_dir, "--train_annotations", tr_ann,
                    "--val_images", va_dir, "--val_annotations", va_ann, "--output_dir", out_dir,
                    "--batch_size", "4", "--val_frequency", "1", "--log_frequency", "1"]
        losses = []
        real_backward = torch.Tensor.backward

        def spy_backward(self, *a, **k):
            losses.append(float(self.detach()))
            return real_backward(self, *a, **k)

        torch.Tensor.backward = spy_backward
        try:
            train.main()
        finally:
            torch.Tensor.backward = real_backward
        from dino_detector import _dod
        final = os.path.join(out_dir, "dino_detector_final.pth")
        ckpt = os.path.join(out_dir, "dino_detector_epoch_2.pth")
        sd = torch.load(final, map_location="cpu")
        ck = torch.load(ckpt, map_location="cpu")
        report.update(losses=losses, launches=int(_dod.launch_count()), final_keys=len(sd),
                      ckpt_has_optimizer="optimizer_state_dict" in ck, ckpt_epoch=ck["epoch"],
                      val_predictions=os.path.exists(os.path.join(out_dir, "val_predictions_epoch_2.json")),
                      val_metrics=os.path.exists(os.path.join(out_dir, "val_metrics_epoch_2.json")))
    else:
        raise SystemExit(f"unknown mode {mode}")
    with open(os.path.join(work, "report.json"), "w") as fh:
        json.dump(report, fh)


if __name__ == "__main__":
    main()
