"""Developer probe: deformable sampling forward at the bench shape (64 images x 50 queries, 8 heads x 2 points x 96)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
b, q, h, p, dh, gh, gw = 64, 50, 8, 2, 96, 10, 137
value = torch.randn(b * gh * gw, h * dh, device="cuda").bfloat16()
raw = torch.randn(b * q, 56, device="cuda")
def f(): return ops.deform_sample(value, raw[:, 48:50], raw[:, :32], raw[:, 32:48], b, q, h, p, dh, gh, gw)
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): f()
e1.record(); torch.cuda.synchronize()
print(f"DOD_DEFORM_VEC={os.environ.get('DOD_DEFORM_VEC', '1')}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
