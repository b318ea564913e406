"""Developer probe: where a C4 train step (L/14 LoRA r=8, deformable decoder) spends its time.
Prints GPU-busy fraction (sum of kernel time / wall) and the top kernels via torch.profiler."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _inputs as synth
from dino_detector.models import DINOv2ObjectDetector
from dino_detector.losses import SetCriterion
from dino_detector.matching import HungarianMatcher

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = DINOv2ObjectDetector(dino_model_name="facebook/dinov2-large", lora_r=8).cuda().train()
crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
x = torch.rand(batch, 3, 518, 518).cuda()
targets = [{k: v.cuda() for k, v in t.items()} for t in synth.make_targets(batch, max_gt=20, seed=3, min_gt=1)]


def step():
    opt.zero_grad()
    loss = sum(crit(m(x), targets).values())
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time_total for e in ev) / 3 / 1e3
print(json.dumps({"batch": batch, "wall_ms_per_step": wall, "gpu_kernel_ms_per_step": tot, "gpu_busy": tot / wall}))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:25]:
    print(f"{e.device_time_total / 3 / 1e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:110]}")
print("kernel launches per step:", sum(e.count for e in ev) // 3)

# ---- per-mode breakdown of the elementwise launches (CUDA events around each call) ----
from dino_detector import ops as _ops
_rec = []
_orig = _ops.eltwise
def _wrapped(mode, a, *args, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = _orig(mode, a, *args, **kw)
    e1.record()
    _rec.append((mode, tuple(a.shape), str(a.dtype), e0, e1))
    return r
_ops.eltwise = _wrapped
step()
torch.cuda.synchronize()
agg = {}
for mode, shape, dt, e0, e1 in _rec:
    k = (mode, shape, dt)
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + e0.elapsed_time(e1))
names = {getattr(_ops, n): n for n in dir(_ops) if n.startswith("ELT_")}
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t:7.3f} ms x{n:3d} {names.get(k[0], k[0])} {k[1]} {k[2]}")
