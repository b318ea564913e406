"""Developer probe: the two residual GEMMs of a B/14 layer (proj, fc2) at batch 64, plain and as folded-LayerNorm
producers, timed alternately."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
m = int(os.environ.get("M", 87680))
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).bfloat16()
def timeit(f, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
d = 768
h16 = torch.empty((m, d), dtype=torch.bfloat16, device="cuda")
stats = torch.zeros((2 * ((d + 255) // 256), m, 2), device="cuda")
for name, (n, k) in {"proj": (768, 768), "fc2": (768, 3072)}.items():
    a, w = rnd(m, k), rnd(n, k)
    bias = torch.randn(n, device="cuda") * 0.1
    scale = torch.ones(n, device="cuda")
    r = torch.randn(m, n, device="cuda")
    out = torch.empty((m, n), dtype=torch.float32, device="cuda")
    plain = lambda: ops.gemm(a, w, bias, scale=scale, residual=r, out=out)
    fold = lambda: ops.gemm(a, w, bias, scale=scale, residual=r, out=out, ln_out=(h16, stats))
    inplace = lambda: ops.gemm(a, w, bias, scale=scale, residual=out, out=out, ln_out=(h16, stats))
    for _ in range(3): plain(); fold(); inplace()
    tp, tf, ti = [], [], []
    for _ in range(5):
        tp.append(timeit(plain)); tf.append(timeit(fold)); ti.append(timeit(inplace))
    tp.sort(); tf.sort(); ti.sort()
    print(f"{name:5s} plain {tp[2]:7.1f} us   producer {tf[2]:7.1f} us   producer in place {ti[2]:7.1f} us", flush=True)
