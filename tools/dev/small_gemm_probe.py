"""Developer probe: the decoder's GEMM shapes (M = 64 x 50 query rows) under the pair kernel and the single-CTA kernel
(DOD_GEMM_PAIR_MIN_M), GPU time from a CUDA graph of 20 calls."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
m = int(os.environ.get("M", 3200))
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).bfloat16()
def gtime(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        f()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(it): f()
    torch.cuda.synchronize(); gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
tot = 0
for name, (n, k, res, act) in {"sa_in": (2304, 768, False, ops.ACT_NONE), "sa_out": (768, 768, True, ops.ACT_NONE),
                               "ca_out": (768, 768, True, ops.ACT_NONE), "l1": (1024, 768, False, ops.ACT_RELU),
                               "l2": (768, 1024, True, ops.ACT_NONE), "box0": (384, 768, False, ops.ACT_RELU)}.items():
    a, w = rnd(m, k), rnd(n, k)
    bias = torch.randn(n, device="cuda") * 0.1
    r = torch.randn(m, n, device="cuda") if res else None
    out = torch.empty((m, n), dtype=torch.float32 if res else torch.bfloat16, device="cuda")
    t = gtime(lambda: ops.gemm(a, w, bias, act=act, residual=r, out=out))
    tot += t
    print(f"{name:7s} N={n:4d} K={k:4d}: {t:6.1f} us", flush=True)
print(f"PAIR_MIN_M={os.environ.get('DOD_GEMM_PAIR_MIN_M', '512')} M={m}: sum {tot:.1f} us")
