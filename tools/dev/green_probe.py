"""Developer probe: two green contexts (disjoint SM partitions), attention on one, the layer's GEMMs on the other,
alone and concurrently."""
import os, sys, time
import torch
from cuda.bindings import driver as cu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))

def ck(r):
    err, *rest = r
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"driver error {err}")
    return rest[0] if len(rest) == 1 else rest

n_attn = int(sys.argv[1]) if len(sys.argv) > 1 else 56
torch.zeros(1, device="cuda")
dev = ck(cu.cuDeviceGet(0))
res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs", res.sm.smCount)
groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, n_attn))
print("split:", nb, "group(s) of", groups[0].sm.smCount, "remaining", rem.sm.smCount)
streams = []
for r in (groups[0], rem):
    desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
    g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    streams.append((g, s, r.sm.smCount))
sA = torch.cuda.ExternalStream(int(streams[0][1]))
sG = torch.cuda.ExternalStream(int(streams[1][1]))
smA, smG = streams[0][2], streams[1][2]
os.environ["DOD_GEMM_MAX_PAIRS"] = str(smG // 2)
from dino_detector import ops

b, s, h = 32, 1370, 12
d = h * 64
m = b * s
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*sh): return (torch.randn(*sh, device="cuda", generator=g) * 0.05).bfloat16()
qkv = (torch.randn(m, 3 * d, device="cuda") * 0.5).bfloat16()
ctx = torch.empty(m, d, dtype=torch.bfloat16, device="cuda")
gemms = []
for name, (n, k, r, act) in {"proj": (768, 768, True, ops.ACT_NONE), "fc1": (3072, 768, False, ops.ACT_GELU_ERF),
                             "fc2": (768, 3072, True, ops.ACT_NONE), "qkv": (2304, 768, False, ops.ACT_NONE)}.items():
    a, w = rnd(m, k), rnd(n, k)
    bias = torch.randn(n, device="cuda") * 0.1
    scale = torch.ones(n, device="cuda") if r else None
    rs = torch.randn(m, n, device="cuda") if r else None
    out = torch.empty((m, n), dtype=torch.float32 if r else torch.bfloat16, device="cuda")
    gemms.append((a, w, bias, act, scale, rs, out))
def run_attn():
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125, out=ctx)
def run_gemms():
    for a, w, bias, act, scale, rs, out in gemms:
        ops.gemm(a, w, bias, act=act, scale=scale, residual=rs, out=out)
torch.cuda.synchronize()

def timed(fa, fg, it=8):
    for _ in range(2):
        if fa:
            with torch.cuda.stream(sA): fa()
        if fg:
            with torch.cuda.stream(sG): fg()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        if fa:
            with torch.cuda.stream(sA): fa()
        if fg:
            with torch.cuda.stream(sG): fg()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it * 1e6

# reference: everything on the whole chip, sequentially (default stream, no cap)
os.environ.pop("DOD_GEMM_MAX_PAIRS")
def seq():
    run_attn(); run_gemms()
for _ in range(2): seq()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): seq()
torch.cuda.synchronize()
t_seq = (time.perf_counter() - t0) / 8 * 1e6
os.environ["DOD_GEMM_MAX_PAIRS"] = str(smG // 2)
ta = timed(run_attn, None)
tg = timed(None, run_gemms)
tb = timed(run_attn, run_gemms)
print(f"half batch (32 images), one layer: whole chip sequential {t_seq:.0f} us | attention alone on {smA} SMs {ta:.0f} us | "
      f"GEMM chain alone on {smG} SMs {tg:.0f} us | both concurrently {tb:.0f} us")
