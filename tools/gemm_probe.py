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

# LayerNorm folded across two GEMMs: producer (residual GEMM + bf16 copy + row partial sums) and consumer
# (rstd scaling in the epilogue) variants of the same shapes, timed ALTERNATELY with the plain form (the
# chip's clock under the power cap drifts over a run, so back-to-back blocks are not comparable)
d = 768
h16 = torch.empty((m, d), dtype=torch.bfloat16, device="cuda")
stats = torch.zeros((2 * ((d + 255) // 256), m, 2), device="cuda")
stats[:, :, 1] = 128.0   # var = 1, rstd = 1
rstd = ops.ln_rstd(stats, d, 1e-6)


def timeit(f, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, (n, k, res, act) in shapes.items():
    a, w = rnd(m, k), rnd(n, k)
    bias = torch.randn(n, device="cuda") * 0.1
    scale = torch.ones(n, device="cuda") if res else None
    r = torch.randn(m, n, device="cuda") if res else None
    out = torch.empty((m, n), dtype=torch.float32 if res else torch.bfloat16, device="cuda")
    plain = lambda: ops.gemm(a, w, bias, act=act, scale=scale, residual=r, out=out)
    if res:
        fold = lambda: ops.gemm(a, w, bias, act=act, scale=scale, residual=r, out=out, ln_out=(h16, stats))
    else:
        fold = lambda: ops.gemm(a, w, bias, act=act, out=out, row_scale=rstd)
    for _ in range(3): plain(); fold()
    tp, tf = [], []
    for _ in range(5):
        tp.append(timeit(plain)); tf.append(timeit(fold))
    tp.sort(); tf.sort()
    print(f"{name:5s} plain {tp[2]:7.1f} us (min {tp[0]:.1f})   folded-LN {'producer' if res else 'consumer'} {tf[2]:7.1f} us (min {tf[0]:.1f})")
if "--no-cublas" in sys.argv: sys.exit(0)

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
