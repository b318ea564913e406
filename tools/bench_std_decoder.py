"""Developer probe: BASELINE configs[1] with the STANDARD decoder (`use_deformable=False`, 100 queries): B/14, bf16,
64 x 518x518, next to the default (deformable) constructor.  Prints one JSON line.
    python tools/bench_std_decoder.py [--steps 10]"""
import argparse, contextlib, io, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import _dod, ops
from dino_detector.models import DINOv2ObjectDetector

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
out = {}
x = torch.rand(a.batch, 3, 518, 518, device="cuda")
for name, kw in (("default_deformable_50q", {}), ("standard_100q", dict(use_deformable=False, num_queries=100))):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()), torch.device("cuda"):
        m = DINOv2ObjectDetector(dino_model_name="facebook/dinov2-base", **kw)
    m.precision = "bf16"
    m = m.cuda().eval()
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        ops.profile_begin()
        _dod.launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        prof = ops.profile_end()
    ms = e0.elapsed_time(e1) / a.steps
    # decoder alone: memory of the last forward through the decoder pack
    mem, b, n = m.backbone.forward_rows(x)
    with torch.no_grad():
        for _ in range(2):
            m.decoder.forward_rows(mem, b, n)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            m.decoder.forward_rows(mem, b, n)
        e1.record()
        torch.cuda.synchronize()
    out[name] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3, "decoder_ms": e0.elapsed_time(e1) / a.steps,
                 "launches_per_step": _dod.launch_count() / a.steps}
    del m
    torch.cuda.empty_cache()
print(json.dumps(out))
