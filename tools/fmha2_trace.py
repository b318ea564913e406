"""Developer probe: SM-clock stamps of one persistent fmha2 CTA (lib built with -DDOD_FMHA_TRACE:
tools/build_variant.sh trace attention.cu -DDOD_FMHA_TRACE; DOD_LIB=.../libdod_trace.so python tools/fmha2_trace.py)."""
import ctypes, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops, _dod
b, s, h = 64, 1370, 12
d = h * 64
qkv = (torch.randn(b * s, 3 * d, device="cuda") * 0.5).bfloat16()
for _ in range(2):
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
torch.cuda.synchronize()
lib = ctypes.CDLL(_dod.LIB_PATH)
T0, N, P = 22, 14, 12
buf = (ctypes.c_longlong * (4 * N * P))()
assert lib.dod_debug_fmha2_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(4, N, P)
t0 = t[t > 0].min()
names = {0: ["top", "s_full", "exp0", "pend", "exp1", "ldwait", "max", "exp2", "post", "exp3", "st/l", "peer"],
         2: ["top", "k_full", "pv_done", "issued"], 3: ["top", "v_full", "p_full", "issued"]}
names[1] = names[0]
for slot, label in ((0, "softmax warp0 (half 0)"), (1, "softmax warp4 (half 1)"), (2, "QK thread"), (3, "PV thread")):
    print(label, names[slot])
    for j in range(N):
        row = t[slot, j, :len(names[slot])]
        print(f"  t={T0 + j:3d} " + " ".join(f"{(x - t0) if x > 0 else -1:7d}" for x in row))
