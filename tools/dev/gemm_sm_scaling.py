"""Developer probe: the encoder GEMMs of a B/14 layer on a SUBSET of the SMs (DOD_GEMM_MAX_PAIRS): does the per-SM
rate rise when fewer SMs share the L2 bandwidth?  Prints us and the per-SM efficiency relative to the full chip."""
import os, subprocess, sys, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "dinov2-od_b200"))
    from dino_detector import ops
    m = int(sys.argv[2])
    g = torch.Generator(device="cuda").manual_seed(0)
    def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).bfloat16()
    res = {}
    for name, (n, k, r, act) in {"qkv": (2304, 768, False, ops.ACT_NONE), "proj": (768, 768, True, ops.ACT_NONE),
                                 "fc1": (3072, 768, False, ops.ACT_GELU_ERF), "fc2": (768, 3072, True, ops.ACT_NONE)}.items():
        a, w = rnd(m, k), rnd(n, k)
        bias = torch.randn(n, device="cuda") * 0.1
        scale = torch.ones(n, device="cuda") if r else None
        rs = torch.randn(m, n, device="cuda") if r else None
        out = torch.empty((m, n), dtype=torch.float32 if r else torch.bfloat16, device="cuda")
        f = lambda: ops.gemm(a, w, bias, act=act, scale=scale, residual=rs, out=out)
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20 * 1e3
    print(json.dumps(res))
    sys.exit(0)
m = 87680
base = None
for pairs in (74, 64, 56, 48, 42, 37):
    env = dict(os.environ, DOD_GEMM_MAX_PAIRS=str(pairs))
    out = subprocess.run([sys.executable, __file__, "child", str(m)], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    r = json.loads(out)
    if base is None: base = r
    print(f"pairs {pairs:3d} ({2*pairs:3d} SMs): " + "  ".join(f"{k} {v:7.1f} us (per-SM rate x{base[k] * 74 / (v * pairs):.3f})" for k, v in r.items()), flush=True)
