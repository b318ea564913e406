"""Developer probe: LoRA weight gradients of one L/14 block at batch 32 -- dod_lowrank_wgrad vs the batched GEMM."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
from dino_detector import ops
b, n = 32, 1370
m = b * n
def timeit(f, it=10):
    """GPU time of f: `it` calls captured in one CUDA graph (the host side of a ctypes launch is ~10-30 us)."""
    for _ in range(2): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f()
        with torch.cuda.graph(g, stream=s):
            for _ in range(it): f()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
def timeit_eager(f, it=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
tot_old = tot_new = 0.0
for name, cols, blocks, r in [("qkv dB", 3072, 3, 8), ("qkv dA", 1024, 1, 24), ("proj dB", 1024, 1, 8), ("proj dA", 1024, 1, 8),
                              ("fc1 dB", 4096, 1, 8), ("fc1 dA", 1024, 1, 8), ("fc2 dB", 1024, 1, 8), ("fc2 dA", 4096, 1, 8)]:
    big = torch.randn(m, cols, device="cuda").bfloat16()
    small = torch.zeros(m, 64, device="cuda", dtype=torch.bfloat16)
    small[:, :blocks * r] = torch.randn(m, blocks * r, device="cuda").bfloat16()
    tr = name.endswith("dA")
    o_old = torch.zeros((blocks * r, cols) if tr else (cols, blocks * r), device="cuda")
    o_new = torch.zeros((blocks, r, cols // blocks) if tr else (blocks, cols // blocks, r), device="cuda")
    scratch = torch.empty(b * cols * r, device="cuda")
    t_old = timeit(lambda: ops.lowrank_wgrad(big, small, blocks * r, o_old, transposed=tr))
    t_new = timeit(lambda: ops.lowrank_wgrad_tc(big, small, r, o_new, transposed=tr, splits=b, blocks=blocks, scratch=scratch))
    tot_old += t_old; tot_new += t_new
    print(f"{name:8s} cols {cols:4d} r {blocks}x{r:2d}: dod_lowrank_wgrad {t_old:7.1f} us   batched GEMM + colsum {t_new:7.1f} us "
          f"({big.numel() * 2 / t_new / 1e6:.2f} TB/s of the big operand)", flush=True)
print(f"per block: {tot_old:.0f} us -> {tot_new:.0f} us")
