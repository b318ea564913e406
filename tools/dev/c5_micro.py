"""Developer probe: g/14 inference throughput vs micro-batch (bench.measure_inference)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
cx = bench.Ctx()
for micro in (32, 64, 128):
    r = bench.measure_inference(cx, "facebook/dinov2-giant", 128, micro, 2, 1, "giant")
    print(micro, round(r["value"], 1), "img/s", round(r["ms_per_step"], 1), "ms", round(r["frac_of_sustained_bf16_peak"], 3), flush=True)
for micro in (32, 64):
    r = bench.measure_inference(cx, "facebook/dinov2-large", 128, micro, 3, 2, "large")
    print("L/14", micro, round(r["value"], 1), "img/s", round(r["frac_of_sustained_bf16_peak"], 3), flush=True)
