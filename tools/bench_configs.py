#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (the contract bench is bench.py).

    python tools/bench_configs.py [c1] [c3] [c4] [c5]      # one JSON line per config

c1  lightweight S/14 + decoder (100 queries), fp32 mode, batch 2 at 224x224: latency
c3  Hungarian matcher, batch 256, 100 queries x <= 50 GT: us/batch on the GPU (the CPU-oracle timing of the same batch is test_matcher_gpu / profiles/r01_configs.jsonl)
c4  L/14 LoRA r=8 + decoder train step (forward + SetCriterion + backward + Adam), bf16, one GPU
c5  g/14 detector inference, bf16, 518x518, micro-batches of 32
All timings: CUDA events on the launching stream, warm-up first; synthetic data, random-init weights.
"""
import contextlib
import io
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _timed(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _model(**kw):
    from dino_detector.models import DINOv2ObjectDetector
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DINOv2ObjectDetector(**kw)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "lora_B" in name or "sampling_offsets" in name or "attention_weights" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.cuda()


def c1():
    m = _model(dino_model_name="facebook/dinov2-small", hidden_dim=256, num_queries=100, num_decoder_layers=2,
               dim_feedforward=512, lora_r=1, nheads=4).eval()
    x = torch.rand(2, 3, 224, 224).cuda()
    out = {}
    for prec in ("fp32", "bf16"):
        m.precision = prec
        with torch.no_grad():
            out[prec + "_ms"] = _timed(lambda: m(x), 20, 3)
    from dino_detector.runtime import GraphedDetector
    m.precision = "bf16"
    graphed = GraphedDetector(m, x)
    out["bf16_cuda_graph_ms"] = _timed(lambda: graphed(x), 50, 5)
    print(json.dumps({"config": "c1 lightweight S/14 + deformable decoder (100 q), batch 2, 224x224", **out}))


def c3():
    import _inputs as synth
    from dino_detector.matching import HungarianMatcher
    preds = synth.make_predictions(256, 100, seed=0)
    targets = synth.make_targets(256, max_gt=50, seed=0)
    dev_preds = {k: v.cuda() for k, v in preds.items()}
    dev_targets = [{k: v.cuda() for k, v in t.items()} for t in targets]
    m = HungarianMatcher()
    packed = m.pack_targets(dev_targets, torch.device("cuda"))
    kern_ms = _timed(lambda: m.match_device(dev_preds, dev_targets, packed=packed), 50, 5)
    api_ms = _timed(lambda: m(dev_preds, dev_targets), 20, 3)
    n_pairs = sum(min(100, len(t["labels"])) for t in targets)
    print(json.dumps({"config": "c3 matcher batch 256, 100 queries x <=50 GT", "cost+lsap_kernels_us": 1e3 * kern_ms,
                      "api_incl_target_packing_and_d2h_us": 1e3 * api_ms, "problems_per_s": 256 / (kern_ms * 1e-3),
                      "matched_pairs": n_pairs}))


def c4(batch=8):
    import _inputs as synth
    from dino_detector.losses import SetCriterion
    from dino_detector.matching import HungarianMatcher
    m = _model(dino_model_name="facebook/dinov2-large", lora_r=8).train()
    crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    x = torch.rand(batch, 3, 518, 518).cuda()
    targets = [{k: v.cuda() for k, v in t.items()} for t in synth.make_targets(batch, max_gt=20, seed=3, min_gt=1)]

    def step():
        opt.zero_grad()
        loss = sum(crit(m(x), targets).values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()

    ms = _timed(step, 5, 2)
    m.eval()
    with torch.no_grad():
        fwd = _timed(lambda: m(x), 5, 2)
    print(json.dumps({"config": f"c4 L/14 LoRA r=8 + deformable decoder train step, bf16, batch {batch} at 518x518, 1 GPU",
                      "train_step_ms": ms, "images_per_s": batch / (ms * 1e-3), "inference_forward_ms": fwd,
                      "trainable_params": sum(p.numel() for p in m.parameters() if p.requires_grad)}))


def c5(batch=32):
    m = _model(dino_model_name="facebook/dinov2-giant").eval()
    x = torch.rand(batch, 3, 518, 518).cuda()
    with torch.no_grad():
        ms = _timed(lambda: m(x), 5, 2)
    gf = 3566.69 + 6.28
    print(json.dumps({"config": f"c5 g/14 detector inference bf16, micro-batch {batch} at 518x518, 1 GPU",
                      "ms_per_batch": ms, "images_per_s": batch / (ms * 1e-3),
                      "model_tflops": gf * batch / ms / 1e3}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
    for w in which:
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[w]()
