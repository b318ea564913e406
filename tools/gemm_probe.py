"""Developer probe: the four encoder GEMM shapes of the bench (B/14, 64 x 518x518) timed alone."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops
m = 87680
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).bfloat16()
shapes = {"qkv": (2304, 768, False, ops.ACT_NONE), "proj": (768, 768, True, ops.ACT_NONE),
          "fc1": (3072, 768, False, ops.ACT_GELU_ERF), "fc2": (768, 3072, True, ops.ACT_NONE)}
for name, (n, k, res, act) in shapes.items():
    a, w = rnd(m, k), rnd(n, k)
    bias = torch.randn(n, device="cuda") * 0.1
    scale = torch.ones(n, device="cuda") if res else None
    r = torch.randn(m, n, device="cuda") if res else None
    out = torch.empty((m, n), dtype=torch.float32 if res else torch.bfloat16, device="cuda")
    f = lambda: ops.gemm(a, w, bias, act=act, scale=scale, residual=r, out=out)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:5s} N={n:4d} K={k:4d}: {ms*1e3:7.1f} us  {2*m*n*k/ms/1e9:6.0f} TFLOP/s")

# library bar on the same shapes: torch.matmul (cuBLASLt), bf16 in / bf16 out, no epilogue at all
for name, (n, k, res, act) in shapes.items():
    a, w = rnd(m, k), rnd(n, k)
    f = lambda: torch.matmul(a, w.t())
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"cuBLAS {name:5s} N={n:4d} K={k:4d}: {ms*1e3:7.1f} us  {2*m*n*k/ms/1e9:6.0f} TFLOP/s (plain matmul, bf16 out)")
