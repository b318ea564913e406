"""Developer probe: the attention kernel alone on the bench shape, next to the library bar
(`F.scaled_dot_product_attention`, the call HF `modeling_dinov2.py:215-229` makes: cuDNN / flash / mem-efficient
backends, bf16, [B, H, S, 64]).  Prints one JSON line.

    python tools/fmha_probe.py [--variant base|large|giant] [--batch 64] [--no-sdpa]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="base")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seq", type=int, default=1370)
ap.add_argument("--no-sdpa", action="store_true")
a = ap.parse_args()
h = {"small": 6, "base": 12, "large": 16, "giant": 24}[a.variant]
b, s = a.batch, a.seq
d = h * 64
flops = 4.0 * b * h * s * s * 64


def timeit(f, n=10, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


qkv = (torch.randn(b * s, 3 * d, device="cuda") * 0.5).bfloat16()
out = {"shape": [b, h, s, 64]}
us = sorted(timeit(lambda: ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)) for _ in range(3))
out["libdod_fmha_us"] = us[1]
out["libdod_fmha_tflops"] = flops / us[1] / 1e6
if not a.no_sdpa:
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q, k, v = (qkv.view(b, s, 3, h, 64)[:, :, i].permute(0, 2, 1, 3).contiguous() for i in range(3))
    ref = None
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(be):
                f = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125)  # noqa: E731
                us = sorted(timeit(f) for _ in range(3))
                ref = f()
            out[f"sdpa_{name}_us"] = us[1]
            out[f"sdpa_{name}_tflops"] = flops / us[1] / 1e6
        except Exception as e:  # backend not available for this shape / build
            out[f"sdpa_{name}_us"] = None
            out[f"sdpa_{name}_error"] = f"{type(e).__name__}: {e}"[:160]
    if ref is not None:
        mine = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125).view(b, s, h, 64).permute(0, 2, 1, 3)
        out["max_abs_diff_vs_sdpa"] = float((mine.float() - ref.float()).abs().max())
print(json.dumps(out))
