"""Developer probe: im2col kernel alone on the bench shape (64 x 3 x 518 x 518 fp32 / uint8 -> [87616, 592] bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops
b = 64
xf = torch.rand(b, 3, 518, 518, device="cuda")
xu = (torch.rand(b, 518, 518, 3, device="cuda") * 255).to(torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, x in (("fp32 NCHW", xf), ("uint8 NHWC", xu)):
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.patchify14(x, 592); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    byts = x.numel() * x.element_size() + b * 1369 * 592 * 2
    print(f"patchify {name}: {ts[len(ts)//2]:.1f} us  {byts / ts[len(ts)//2] / 1e3:.0f} GB/s (L2 flushed)")
