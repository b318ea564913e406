"""Developer probe for ncu: the fc1 GEMM (N=3072, K=768, GELU, 87 680 rows) once plain and once as the
consumer of a folded LayerNorm.  ncu -k regex:gemm2 -s 4 -c 2 captures exactly those two launches."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops
m, n, k, d = 87680, 3072, 768, 768
g = torch.Generator(device="cuda").manual_seed(0)
a = (torch.randn(m, k, device="cuda", generator=g) * 0.05).bfloat16()
w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16()
bias = torch.randn(n, device="cuda") * 0.1
out = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
stats = torch.zeros((6, m, 2), device="cuda")
stats[:, :, 1] = 128.0
rstd = ops.ln_rstd(stats, d, 1e-6)
plain = lambda: ops.gemm(a, w, bias, act=ops.ACT_GELU_ERF, out=out)
fold = lambda: ops.gemm(a, w, bias, act=ops.ACT_GELU_ERF, out=out, row_scale=rstd)
for _ in range(2): plain(); fold()
torch.cuda.synchronize()
plain(); fold()
torch.cuda.synchronize()
