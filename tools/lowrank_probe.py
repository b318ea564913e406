"""Developer probe: dod_lowrank_wgrad on the L/14 LoRA-block shapes of the C4 train step (batch 32)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops
m = 32 * 1370
for cols, r in [(1024, 8), (3072, 24), (4096, 8)]:
    big = torch.randn(m, cols, device="cuda").bfloat16()
    small = torch.randn(m, 64, device="cuda").bfloat16()
    out = torch.zeros(cols, r, device="cuda")
    f = lambda: ops.lowrank_wgrad(big, small, r, out, transposed=False)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"cols={cols} r={r}: {us:7.1f} us  {m * cols * 2 / us / 1e6:6.2f} TB/s of the big operand")
