"""Developer probe: bench workload (B/14, 64 x 518x518) eager launches vs one CUDA-graph replay.

The difference is what the host-side launch path and inter-kernel gaps cost at this batch size.
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
import bench
from dino_detector.runtime import GraphedDetector

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
x = torch.rand((bench.BATCH, 3, bench.IMG, bench.IMG), device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def timed(f):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_eager = timed(lambda: model(x))
    gd = GraphedDetector(model, x)
    t_graph = timed(lambda: gd.graph.replay())
    t_eager2 = timed(lambda: model(x))
print(f"eager {t_eager:.3f} ms/step, graph replay {t_graph:.3f} ms/step, eager again {t_eager2:.3f} ms/step")
