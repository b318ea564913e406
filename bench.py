#!/usr/bin/env python
"""Throughput benchmark of the detector hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward of the detector over one batch of synthetic images.
Workload at every N: BASELINE.json configs[1] -- DINOv2-B/14 detector (default
constructor: LoRA r=2, deformable decoder, 50 queries), bf16 compute with fp32
accumulation, 64 synthetic 518x518 images per GPU (weak scaling, pure data parallel, no
data-path collective).  Prints ONE JSON line on rank 0.

  value     images/s over all GPUs with the batch already resident in HBM (CUDA events,
            max over ranks).
  e2e       the same metric through the public API `model(images)` with the batch in
            pinned HOST memory: H2D of the images and D2H of pred_logits/pred_boxes are
            inside the timed region, every step.
  roofline  the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of its launches /
            their summed CUDA-event durations, measured inside the timed region, against
            MEASURED_PEAKS.json (sustained bf16 figure, since it is timed inside a long step).
  cpu_baseline  the CPU oracle (oracle/detector_oracle.py, a restatement of the reference
            forward) on the host cores, rank 0 at N=1, on a bounded sample (2 images).

`--impl reference` times the reference's own algorithm on the host CPU: the reference is
pure Python/PyTorch that needs a HF download and pycocotools, neither available on the
GPU box, so this is the oracle port (oracle/), all host threads, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))

METRIC = "detector images/sec at 518px (DINOv2-B/14, bf16)"
UNIT = "images/s"
MODEL_NAME = "facebook/dinov2-base"
BATCH = int(os.environ.get("DOD_BENCH_BATCH", 64))
IMG = 518


WORKLOAD = ("BASELINE configs[1]: DINOv2-B/14 detector (reference default ctor: LoRA r=2 on the last 2 blocks, "
            "deformable decoder, 50 queries, 91 classes) inference at 518x518, 1370 tokens/image")


def _ncu_traffic():
    """DRAM bytes per GEMM launch (read + write) from the committed `ncu --set full` capture of one
    encoder layer (profiles/): average over its four GEMMs; None if the summary is not there."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_final3_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)["gemm"]["avg_dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def _config(world):
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
            "parallelism": f"dp{world}", "weights": "random-init",
            "l2_note": "per-step activations (>= 270 MB per tensor) exceed the 126 MB L2, no explicit flush"}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(tflops=float(p["bf16_tflops_sustained"]), tflops_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured")
    except Exception:
        return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi SM clock / throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.nvml_rows, self._stop = None, [], threading.Event()

    def start(self):
        # NVML in a thread (a sample every ~5 ms: the timed region is a few hundred ms); the nvidia-smi loop
        # (-lms 100, first line after ~0.3 s) is the fallback when pynvml cannot be loaded
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = (pynvml, h)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self):
        pynvml, h = self.nvml
        while not self._stop.is_set():
            try:
                self.nvml_rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                       pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM),
                                       pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20,
                    "hw_thermal_slowdown": 0x40}
            sm = [r[0] for r in self.nvml_rows]
            reasons = sorted(n for n, b in bits.items() if any(r[2] & b for r in self.nvml_rows))
            return {"sm_mhz": statistics.median(sm) if sm else None,
                    "sm_max_mhz": max(r[1] for r in self.nvml_rows) if sm else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml, 5 ms period over the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        # samples under load: the upper half of the observed clocks' power draw is not tracked
        # separately; the median over the timed region is what is reported
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device):
    import contextlib
    import io
    import torch
    from dino_detector.models import DINOv2ObjectDetector
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = DINOv2ObjectDetector(dino_model_name=MODEL_NAME)     # reference default constructor
    # the reference zero-initialises lora_B / sampling_offsets / attention_weights; randomise them so
    # that no term of the forward is a multiply-by-zero (work is identical either way)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "lora_B" in name or "sampling_offsets" in name or "attention_weights" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.precision = "bf16"
    return model.to(device).eval()


def cpu_forward_rate(state_dict, n_images, repeats, threads):
    """images/s of the CPU oracle (restated reference forward) on `n_images` 518x518 images."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import detector_oracle
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    x = torch.rand((n_images, 3, IMG, IMG), generator=g)
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    detector_oracle.detector_forward(sd, x[:1], dino_model_name=MODEL_NAME)      # warm-up
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        detector_oracle.detector_forward(sd, x, dino_model_name=MODEL_NAME)
        times.append(time.perf_counter() - t0)
    return n_images / statistics.median(times), times


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import contextlib
    import io
    import torch
    from dino_detector.models import DINOv2ObjectDetector
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = DINOv2ObjectDetector(dino_model_name=MODEL_NAME)
    threads = os.cpu_count() or 1
    n_img = 2
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import detector_oracle
    torch.set_num_threads(threads)
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    x = torch.rand((n_img, 3, IMG, IMG), generator=torch.Generator().manual_seed(0))
    for _ in range(max(1, min(args.warmup, 2))):
        detector_oracle.detector_forward(sd, x, dino_model_name=MODEL_NAME)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        detector_oracle.detector_forward(sd, x, dino_model_name=MODEL_NAME)
    dt = time.perf_counter() - t0
    value = n_img * steps / dt
    sample = f"{n_img} of {BATCH} images per step (fp32, eval, no_grad), {steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(_config(max(1, args.gpus)), reference_arm=f"CPU oracle port of the reference forward, "
                                                                 f"{n_img}-image sample per step, fp32"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dino_detector import _dod, ops

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the libdod path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    model = build_model(dev)
    g = torch.Generator().manual_seed(100 + rank)
    host = torch.rand((BATCH, 3, IMG, IMG), generator=g).pin_memory()
    x_dev = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput ----------------
    with torch.no_grad():
        for _ in range(args.warmup):
            out = model(x_dev)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        _dod.launch_count_reset()
        ops.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = model(x_dev)
        e1.record()
        barrier()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))
        prof = ops.profile_end()
        launches = _dod.launch_count()
        clocks = sampler.stop() if rank == 0 else None

        # ---------------- end to end through the public API, host buffers ----------------
        out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
        copy_stream = torch.cuda.Stream(dev)
        bufs = [torch.empty_like(x_dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def e2e_steps(n):
            # double-buffered: the H2D copy of batch i+1 (copy stream) overlaps the forward of batch i
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(copy_stream):
                bufs[0].copy_(host, non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(n):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < n:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(freed[nxt])
                        bufs[nxt].copy_(host, non_blocking=True)
                        ready[nxt].record(copy_stream)
                main.wait_event(ready[cur])
                o = model(bufs[cur])
                freed[cur].record(main)
                for k in o:
                    out_host[k].copy_(o[k], non_blocking=True)
            main.synchronize()

        e2e_steps(max(2, min(args.warmup, 3)))
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_steps(args.steps)
        t1.record()
        barrier()
        ms_e2e = max_over_ranks(t0.elapsed_time(t1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    n_img = BATCH * world * args.steps
    value = n_img / (ms_dev * 1e-3)
    e2e_value = n_img / (ms_e2e * 1e-3)
    gl, gf, gt = prof.get("gemm", (0, 0.0, 1e-9))
    fl, ff, ft = prof.get("fmha", (0, 0.0, 1e-9))
    ll, lb, lt = prof.get("layernorm", (0, 0.0, 1e-9))
    gemm_tflops = gf / (gt * 1e-3) / 1e12
    roof = {"kernel": "dod::gemm_kernel (tcgen05/TMEM/TMA)", "bound": "tensor", "achieved": gemm_tflops,
            "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / peaks["tflops"],
            "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
            "traffic": _ncu_traffic(), "launches_per_step": gl / args.steps,
            "share_of_step": gt / (ms_dev if world == 1 else e0.elapsed_time(e1))}
    extra = {
        "fmha": {"kernel": "dod::fmha_kernel", "bound": "tensor", "achieved": ff / (ft * 1e-3) / 1e12,
                 "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ff / (ft * 1e-3) / 1e12 / peaks["tflops"],
                 "share_of_step": ft / e0.elapsed_time(e1)},
        "layernorm": {"kernel": "dod::layernorm_kernel", "bound": "hbm", "achieved": lb / (lt * 1e-3) / 1e9,
                      "peak": peaks["hbm"], "unit": "GB/s", "frac": lb / (lt * 1e-3) / 1e9 / peaks["hbm"],
                      "share_of_step": lt / e0.elapsed_time(e1)},
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        # bounded sample: 8 forwards of 4 images are ~12 s of CPU work on the box's 16 cores
        rate, times = cpu_forward_rate(model.state_dict(), 4, 8, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"4 of {BATCH} images per step, 8 timed forwards (fp32 CPU oracle of the reference), "
                         f"median {statistics.median(times):.2f} s, total {sum(times):.1f} s"}
    h2d = host.numel() * host.element_size()
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(world),
        "roofline": roof, "roofline_other": extra, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "note": "pinned host images, double-buffered H2D on a copy "
                                                             "stream, outputs copied back every step"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line (the JSON result): libraries that write to fd 1 on their own (NCCL prints
    # its version banner there when NCCL_DEBUG=VERSION/INFO) are sent to stderr for the duration of the run
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
