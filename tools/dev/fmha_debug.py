"""Developer probe: per-(image, head, query tile, warp) error map of dod_fmha_fwd against fp32 SDPA."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
for (b, s, h) in [(1, 256, 1), (1, 257, 1), (1, 300, 1), (1, 384, 1), (2, 257, 6), (1, 1370, 2), (40, 1370, 12)]:
    torch.manual_seed(s)
    d = h * 64
    qkv = (torch.randn(b * s, 3 * d, device="cuda")).bfloat16()
    out = ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
    q, k, v = (qkv.float().view(b, s, 3, h, 64).permute(2, 0, 3, 1, 4))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)   # [b, h, s, 64]
    err = (out.float().view(b, s, h, 64).permute(0, 2, 1, 3) - ref).abs().amax(-1)  # [b, h, s]
    scale = ref.abs().max()
    bad = (err / scale > 2e-2).nonzero()
    print((b, s, h), "max rel", float(err.max() / scale), "bad rows", bad.shape[0], flush=True)
    if bad.shape[0]:
        seen = {}
        for bb, hh, r in bad.tolist():
            key = (bb, hh, r // 128, (r % 128) // 32)
            seen[key] = seen.get(key, 0) + 1
        print("  (image, head, q_tile, warp): rows bad ->", list(seen.items())[:24])
