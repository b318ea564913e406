"""The bar on the same box: the backbone the reference calls (HF transformers Dinov2Model, what
`DINOv2Backbone.forward` runs, models/dinov2_backbone.py:58-67) under PyTorch eager
`torch.autocast('cuda', bf16)` -- cuBLASLt GEMMs + SDPA + ATen elementwise -- timed with CUDA events on
the bench workload (B/14, 64 x 518x518).  Library code only: no reference import, no oracle, none of
our kernels.  The backbone is >= 97 % of the detector's FLOPs, so this bounds what the
unmodified reference can reach on one B200 from above.

    python tools/torch_eager_bar.py [--variant base] [--batch 64] [--steps 10]
"""
import argparse
import json

import torch
from transformers import Dinov2Config, Dinov2Model

VARIANTS = {"small": (384, 12, 6, False), "base": (768, 12, 12, False), "large": (1024, 24, 16, False),
            "giant": (1536, 40, 24, True)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="base")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dim, layers, heads, swiglu = VARIANTS[a.variant]
    torch.manual_seed(0)
    cfg = Dinov2Config(image_size=518, patch_size=14, hidden_size=dim, num_hidden_layers=layers,
                       num_attention_heads=heads, use_swiglu_ffn=swiglu)
    model = Dinov2Model(cfg).cuda().eval()
    x = torch.rand(a.batch, 3, 518, 518, device="cuda")
    out = {}
    for mode in ("autocast_bf16", "bf16_weights"):
        m = model if mode == "autocast_bf16" else model.bfloat16()
        xin = x if mode == "autocast_bf16" else x.bfloat16()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
            for _ in range(a.warmup):
                m(xin).last_hidden_state
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                m(xin).last_hidden_state
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[mode] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3}
    print(json.dumps({"what": "HF Dinov2Model backbone only, PyTorch eager on this GPU", "variant": a.variant,
                      "batch": a.batch, "attn": cfg._attn_implementation, "torch": torch.__version__, **out}))


if __name__ == "__main__":
    main()
