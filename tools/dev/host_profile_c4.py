"""Developer probe: where the HOST spends a C4 train step (cProfile over 5 steps; the GPU runs asynchronously)."""
import cProfile, os, pstats, sys, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _inputs as synth
from dino_detector.models import DINOv2ObjectDetector
from dino_detector.losses import SetCriterion
from dino_detector.matching import HungarianMatcher
from dino_detector.optim import FusedAdam
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = DINOv2ObjectDetector(dino_model_name="facebook/dinov2-large", lora_r=8).cuda().train()
crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
crit.strict = False
opt = FusedAdam(m.parameters(), lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
x = torch.rand(batch, 3, 518, 518).cuda()
targets = [{k: v.cuda() for k, v in t.items()} for t in synth.make_targets(batch, max_gt=20, seed=3, min_gt=1)]
def step():
    opt.zero_grad()
    loss = sum(crit(m(x), targets).values())
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): step()
t_host = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5 * 1e3
print(f"host-side issue time {t_host:.1f} ms/step, with the final sync {t_all:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
