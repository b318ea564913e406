"""Developer probe: the four encoder GEMM shapes through torch.matmul (cuBLASLt), for an ncu look at the library's
tile / cluster / shared-memory choices (grid, cluster, dynamic smem, registers, tensor-pipe activity)."""
import torch
m = 87680
g = torch.Generator(device="cuda").manual_seed(0)
for n, k in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.05).bfloat16()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16()
    for _ in range(3):
        c = torch.matmul(a, w.t())
    torch.cuda.synchronize()
print("ok")
