"""Developer probe: time the attention kernel alone (B/14 shapes) on the bench shape."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops
b, s, h = 64, 1370, 12
d = h * 64
qkv = (torch.randn(b * s, 3 * d, device="cuda") * 0.5).bfloat16()
for _ in range(3):
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"fmha B/14 batch 64: {ms*1e3:.1f} us  {4*b*h*s*s*64/ms/1e9:.0f} TFLOP/s-eq")
