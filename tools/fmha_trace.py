"""Developer probe: per-phase SM-clock stamps of one FMHA CTA (needs lib built with -DDOD_FMHA_TRACE:
tools/build_variant.sh trace attention.cu -DDOD_FMHA_TRACE; DOD_LIB=.../libdod_trace.so python tools/fmha_trace.py)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dinov2-od_b200"))
from dino_detector import ops, _dod
b, s, h = 64, 1370, 12
d = h * 64
qkv = (torch.randn(b * s, 3 * d, device="cuda") * 0.5).bfloat16()
for _ in range(3):
    ops.fmha(qkv, b, s, h, q_off=0, k_off=d, v_off=2 * d, scale=0.125)
torch.cuda.synchronize()
lib = ctypes.CDLL(_dod.LIB_PATH)
buf = (ctypes.c_longlong * (3 * 16 * 8))()
assert lib.dod_debug_fmha_trace(buf) == 0
import numpy as np
t = np.array(buf, dtype=np.int64).reshape(3, 16, 8)
t0 = t[t > 0].min()
names = {0: ["top", "s_full", "max", "xchg", "exp", "o_full", "p_stored"], 1: None, 2: ["top", "s_free", "S issued", "p_full(PV warp)", "PV issued", "V refilled"]}
names[1] = names[0]
for slot, label in ((0, "softmax warp0 (half 0)"), (1, "softmax warp4 (half 1)"), (2, "MMA thread")):
    print(label, names[slot])
    for j in range(11):
        row = t[slot, j, :len(names[slot])]
        print(f"  j={j:2d} " + " ".join(f"{(x - t0) if x > 0 else -1:7d}" for x in row))
