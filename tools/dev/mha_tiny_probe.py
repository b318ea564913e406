"""Developer probe: the decoder's query self-attention (64 images x 8 heads x 50 x 50 x 96) -- GPU time from a CUDA graph."""
import os, sys, math, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
b, l, h, dh = 64, 50, 8, 96
d = h * dh
for pad, name in ((0, "16-byte rows (bf16 K/V in shared memory)"), (4, "ld % 8 != 0 (fp32 staging)")):
    buf = (torch.randn(b * l, 3 * d + pad, device="cuda")).bfloat16()
    f = lambda: ops.mha_small(buf[:, :d], buf[:, d:2 * d], buf[:, 2 * d:3 * d], b, l, l, h, dh, 1 / math.sqrt(dh))
    for _ in range(3): f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        f()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(20): f()
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
