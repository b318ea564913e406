#!/usr/bin/env python
"""Throughput benchmark of the detector hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cpu-baseline] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward of the detector over one batch of synthetic images.  Workload of the headline
numbers at every N: BASELINE.json configs[1] -- DINOv2-B/14 detector (default constructor: LoRA r=2,
deformable decoder, 50 queries), bf16 compute with fp32 accumulation, 64 synthetic 518x518 images per GPU
(weak scaling, pure data parallel, no data-path collective).  Prints ONE JSON line on rank 0.

  value     images/s over all GPUs with the batch already resident in HBM (CUDA events, max over ranks).
  e2e       the same metric through the public API `model(images)` with the batch in pinned HOST memory: H2D of
            the images and D2H of pred_logits/pred_boxes are inside the timed region, every step.
  roofline  the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of its launches / their summed CUDA-event
            durations, measured inside the timed region, against MEASURED_PEAKS.json (sustained bf16 figure,
            since it is timed inside a long step; the burst fraction is given next to it).
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, reference models/detector.py:58-69) on the host cores,
            rank 0 at N=1, on a bounded sample (2 images per forward).
  records   the other BASELINE.json configs, measured at this N in the same run (each: images/s of the whole job,
            ms/step, model-level TFLOP/s and its fraction of the sustained / burst bf16 peak):
              infer_l14   L/14 detector inference, 64 images per GPU          (north_star's target model)
              infer_c5    g/14 detector inference, global batch 512 split 512/N per GPU, micro-batches of 64
              train_c4    L/14 LoRA r=8 + deformable decoder FULL train step (forward, GPU matcher, fused
                          criterion, hand-written backward, ONE NCCL all-reduce of the flat gradient inside the
                          timed step, global-norm clip + Adam), 32 images per GPU
              ddp_parity  (N >= 2) gradients after the all-reduce vs the single-process gradients on the
                          concatenated batch, default switches (LayerNorm fold on): worst_rel over parameters
              torch_eager_b200  (N = 1) the reference module itself (baseline/_ref, standard decoder) on this
                          GPU under torch.autocast(bf16): the same-box bar
              e2e_uint8   e2e through the uint8 [B, H, W, 3] input path (4x fewer H2D bytes)

`--impl reference` times the UNMODIFIED reference from baseline/_ref (its own DINOv2ObjectDetector.forward, default
constructor = deformable decoder with the python sampling loop) on the host CPU with all host threads, on a bounded
sample of the same workload (2 of the 64 images per step); rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))

METRIC = "detector images/sec at 518px (DINOv2-B/14, bf16)"
UNIT = "images/s"
MODEL_NAME = "facebook/dinov2-base"
BATCH = int(os.environ.get("DOD_BENCH_BATCH", 64))
IMG = 518
N_TOK = (IMG // 14) ** 2 + 1
REF_SAMPLE_IMAGES = 2

WORKLOAD = ("BASELINE configs[1]: DINOv2-B/14 detector (reference default ctor: LoRA r=2 on the last 2 blocks, "
            "deformable decoder, 50 queries, 91 classes) inference at 518x518, 1370 tokens/image")

VARIANTS = {"small": (384, 12, 6), "base": (768, 12, 12), "large": (1024, 24, 16), "giant": (1536, 40, 24)}


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


def _ncu_traffic():
    """DRAM bytes per GEMM launch (read + write) from the committed `ncu --set full` capture of one encoder
    layer (profiles/): average over its four GEMMs; None if no summary is there."""
    for name in ("r02_traffic.json", "r01_final3_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                return json.load(fh)["gemm"]["avg_dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            continue
    return None


# ---------------------------------------------------------------------------------------------------
# algorithmic work (SURVEY.md 8d / BASELINE.md section 3)
# ---------------------------------------------------------------------------------------------------
def forward_gflops(variant, hidden_dim=768, n=N_TOK, queries=50, dec_layers=3, ffn=1024, classes=91, heads=8,
                   points=2):
    """GF per image of the detector forward with the default (deformable) decoder: encoder
    L.(24 N D^2 + 4 N^2 D) + patch embed 2 (N-1) 588 D + projection 2 N D Dd (when D != Dd) + decoder
    (value_proj over memory counted ONCE: the reference stacks one shared layer, deformable_attention.py:284)."""
    d, layers, _ = VARIANTS[variant]
    enc_gemm = layers * 24.0 * n * d * d
    enc_attn = layers * 4.0 * n * n * d
    patch = 2.0 * (n - 1) * 588 * d
    dd = hidden_dim
    proj = 2.0 * n * d * dd if d != dd else 0.0
    hp = heads * points
    dec = 2.0 * n * dd * dd                                     # value_proj(memory), layer-invariant
    per_layer = (8.0 * queries * dd * dd + 4.0 * queries * queries * dd       # self-attention
                 + 2.0 * queries * dd * (3 * hp + 2) + 2.0 * queries * dd * dd  # offsets/weights/ref + output_proj
                 + 4.0 * queries * dd * ffn)                                 # FFN
    dec += dec_layers * per_layer + 2.0 * queries * dd * (classes + dd // 2) + 4.0 * queries * dd
    return dict(total=(enc_gemm + enc_attn + patch + proj + dec) / 1e9, enc_gemm=enc_gemm / 1e9,
                enc_attn=enc_attn / 1e9, patch=patch / 1e9, proj=proj / 1e9, decoder=dec / 1e9)


def train_gflops(variant, hidden_dim=768, n=N_TOK, **kw):
    """GF per image of the C4 train step: forward + backward of the decoder and the projection
    (dgrad + wgrad = 2x their forward) + backward of the two LoRA-wrapped encoder layers (dgrad through the frozen
    weights = 1x their GEMM flops; attention backward = 2.5x its forward; LoRA wgrad < 0.1 %)."""
    f = forward_gflops(variant, hidden_dim, n, **kw)
    d, layers, _ = VARIANTS[variant]
    two = 2.0 * (24.0 * n * d * d + 2.5 * 4.0 * n * n * d) / 1e9
    return f["total"] + 2.0 * (f["decoder"] + f["proj"]) + two


# ---------------------------------------------------------------------------------------------------
# clocks: sampled by a SEPARATE process (a thread of this one is starved by the launch loop's GIL)
# ---------------------------------------------------------------------------------------------------
_POLLER = r"""
import sys, time
idx, path = int(sys.argv[1]), sys.argv[2]
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(idx)
mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
with open(path, "w", buffering=1) as fh:
    fh.write("ready %d\n" % mx)
    while True:
        try:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
            fh.write("%.6f %d %d %.1f\n" % (time.time(), sm, rs, pw))
        except Exception as e:
            fh.write("err %r\n" % (e,))
        time.sleep(0.004)
"""


class ClockSampler:
    """NVML SM clock / throttle reasons / power, polled every ~4 ms by a child process into a file; windows of
    interest are cut out afterwards by wall-clock time stamps."""
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, local_index):
        self.proc, self.path = None, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.index = local_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_index < len(ids) and ids[local_index].isdigit():
                self.index = int(ids[local_index])

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="dod_clocks_", suffix=".txt")
            os.close(fd)
            self.proc = subprocess.Popen([sys.executable, "-c", _POLLER, str(self.index), self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 10.0:                      # wait for the first line (NVML initialised)
                with open(self.path) as fh:
                    if fh.readline().startswith("ready"):
                        return
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def window(self, t_start, t_stop):
        """Summary of the samples with t_start <= t <= t_stop."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml poller unavailable"], "samples": 0}
        sm, rs, pw, mx = [], 0, [], None
        try:
            with open(self.path) as fh:
                for line in fh:
                    f = line.split()
                    if f and f[0] == "ready":
                        mx = int(f[1])
                    elif len(f) == 4 and f[0][0].isdigit() and t_start <= float(f[0]) <= t_stop:
                        sm.append(int(f[1]))
                        rs |= int(f[2])
                        pw.append(float(f[3]))
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": mx, "reasons": sorted(n for n, b in self.BITS.items() if rs & b),
                "power_w_median": statistics.median(pw) if pw else None, "samples": len(sm),
                "source": "NVML polled every ~4 ms by a child process over the timed region"}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.path:
            try:
                os.unlink(self.path)
            except OSError:
                pass


# ---------------------------------------------------------------------------------------------------
# models / inputs
# ---------------------------------------------------------------------------------------------------
def build_model(device, name=MODEL_NAME, **kw):
    """Our detector with the reference default constructor (+ overrides), random-init on `device`."""
    import contextlib
    import io
    import torch
    from dino_detector.models import DINOv2ObjectDetector
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()), torch.device(device):
        model = DINOv2ObjectDetector(dino_model_name=name, **kw)
    # the reference zero-initialises lora_B / sampling_offsets / attention_weights; randomise them so
    # that no term of the forward is a multiply-by-zero (work is identical either way)
    g = torch.Generator(device=device).manual_seed(1)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if "lora_B" in pname or "sampling_offsets" in pname or "attention_weights" in pname:
                p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.02)
    model.precision = "bf16"
    return model.to(device).eval()


def make_targets(batch, *, max_gt=20, num_classes=91, seed=0, min_gt=1):
    """COCO-style synthetic targets (dataset.py:102-111): labels int64 [n], boxes cxcywh fp32 [n, 4]."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for _ in range(batch):
        n = int(torch.randint(min_gt, max_gt + 1, (1,), generator=g))
        cxcy = torch.rand((n, 2), generator=g) * 0.6 + 0.2
        wh = torch.rand((n, 2), generator=g) * 0.3 + 0.02
        out.append({"labels": torch.randint(0, num_classes, (n,), generator=g), "boxes": torch.cat([cxcy, wh], dim=1)})
    return out


def reference_forward_rate(n_images, steps, warmup, threads):
    """images/s of the UNMODIFIED reference detector (baseline/_ref) on the host CPU, fp32, eval, no_grad."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = ref_loader.build_detector(dino_model_name=MODEL_NAME).eval()          # reference default constructor
    x = torch.rand((n_images, 3, IMG, IMG), generator=torch.Generator().manual_seed(0))
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            model(x)
        for _ in range(steps):
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
    return n_images * len(times) / sum(times), times


def run_reference(args):
    """--impl reference: the reference's own DINOv2ObjectDetector.forward on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", args.gpus)))
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    if not ref_loader.available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing (pip install --target "
                          "baseline/_ref of the reference was not run in the build container)"}),
              file=_RESULT_OUT, flush=True)
        return
    steps, warmup = max(1, min(args.steps, 30)), max(0, min(args.warmup, 3))     # ~2-4 s per forward
    t_wall = time.perf_counter()
    value, times = reference_forward_rate(REF_SAMPLE_IMAGES, steps, warmup, threads)
    sample = (f"{REF_SAMPLE_IMAGES} of the {BATCH} images of a step per forward, unmodified reference "
              f"DINOv2ObjectDetector (baseline/_ref, default ctor: deformable decoder with its python sampling "
              f"loop), fp32, eval, no_grad, {threads} threads, {steps} timed forwards after {warmup} warm-ups, "
              f"median {statistics.median(times):.2f} s per forward, {time.perf_counter() - t_wall:.0f} s wall")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / collective helpers shared by the measurements."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.rank = int(os.environ.get("RANK", 0))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the libdod path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = _peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """ms for `steps` calls of fn after `warmup` calls: barrier + synchronize on both sides, CUDA events on the
        launching stream, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        t1 = time.time()
        return self.max_over_ranks(e0.elapsed_time(e1)), (t0, t1)

    def rate_record(self, images_per_step_global, ms_total, steps, gflops_per_image, **extra):
        ms = ms_total / steps
        tf = gflops_per_image * images_per_step_global / self.world / ms             # GF / ms = TFLOP/s, per GPU
        rec = {"value": images_per_step_global / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
               "n_gpus": self.world, "images_per_gpu_per_step": images_per_step_global // self.world,
               "gflops_per_image": round(gflops_per_image, 2), "model_tflops_per_gpu": tf,
               "frac_of_sustained_bf16_peak": tf / self.peaks["tflops"],
               "frac_of_burst_bf16_peak": tf / self.peaks["tflops_burst"]}
        rec.update(extra)
        return rec


def measure_headline(cx, args, sampler):
    """configs[1]: device-resident value, e2e (fp32 and uint8 host inputs), per-kernel roofline figures."""
    torch = cx.torch
    from dino_detector import _dod, ops
    dev = cx.dev
    model = build_model(dev)
    g = torch.Generator().manual_seed(100 + cx.rank)
    host = torch.rand((BATCH, 3, IMG, IMG), generator=g).pin_memory()
    host_u8 = (host.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
    x_dev = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    res = {}
    with torch.no_grad():
        for _ in range(args.warmup):
            out = model(x_dev)
        cx.barrier()
        _dod.launch_count_reset()
        ops.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start = time.time()
        e0.record()
        for _ in range(args.steps):
            out = model(x_dev)
        e1.record()
        cx.barrier()
        t_stop = time.time()
        ms_local = e0.elapsed_time(e1)
        res["ms_dev"] = cx.max_over_ranks(ms_local)
        res["ms_local"] = ms_local
        res["prof"] = ops.profile_end()
        res["launches"] = _dod.launch_count()
        res["clocks"] = sampler.window(t_start, t_stop) if sampler is not None else None

        # ---------------- end to end through the public API, host buffers ----------------
        out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
        copy_stream = torch.cuda.Stream(dev)

        def e2e_runner(src):
            bufs = [torch.empty(src.shape, dtype=src.dtype, device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            freed = [torch.cuda.Event() for _ in range(2)]

            def run(n):
                # double-buffered: the H2D copy of batch i+1 (copy stream) overlaps the forward of batch i
                main = torch.cuda.current_stream(dev)
                with torch.cuda.stream(copy_stream):
                    bufs[0].copy_(src, non_blocking=True)
                    ready[0].record(copy_stream)
                for i in range(n):
                    cur, nxt = i & 1, (i + 1) & 1
                    if i + 1 < n:
                        with torch.cuda.stream(copy_stream):
                            if i >= 1:
                                copy_stream.wait_event(freed[nxt])
                            bufs[nxt].copy_(src, non_blocking=True)
                            ready[nxt].record(copy_stream)
                    main.wait_event(ready[cur])
                    o = model(bufs[cur])
                    freed[cur].record(main)
                    for k in o:
                        out_host[k].copy_(o[k], non_blocking=True)
                main.synchronize()
            return run

        for key, src in (("e2e", host), ("e2e_uint8", host_u8)):
            run = e2e_runner(src)
            run(max(2, min(args.warmup, 3)))
            cx.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            run(args.steps)
            t1.record()
            cx.barrier()
            res[key] = dict(ms=cx.max_over_ranks(t0.elapsed_time(t1)), h2d=src.numel() * src.element_size(),
                            d2h=sum(v.numel() * v.element_size() for v in out_host.values()))
    del model, x_dev
    torch.cuda.empty_cache()
    return res


def measure_inference(cx, name, images_per_gpu, micro, steps, warmup, variant):
    """Device-resident inference throughput of another model size (random init on the device)."""
    torch = cx.torch
    model = build_model(cx.dev, name)
    g = torch.Generator(device=cx.dev).manual_seed(200 + cx.rank)
    x = torch.rand((images_per_gpu, 3, IMG, IMG), generator=g, device=cx.dev)
    chunks = [x[i:i + micro] for i in range(0, images_per_gpu, micro)]

    def step():
        for c in chunks:
            model(c)

    with torch.no_grad():
        ms, _ = cx.timed(step, steps, warmup)
    hidden = model.decoder.hidden_dim
    gf = forward_gflops(variant, hidden)["total"]
    del model, x, chunks
    torch.cuda.empty_cache()
    return cx.rate_record(images_per_gpu * cx.world, ms, steps, gf, warmup_steps=warmup, micro_batch=micro,
                          model=name, decoder_hidden_dim=hidden)


def measure_train(cx, steps, warmup, batch=32, parity_images=2):
    """C4: full train step with the flat-gradient all-reduce inside the timed step; at N >= 2 first the gradient
    parity check (after-all-reduce gradients vs single-process gradients on the concatenated batch)."""
    torch, dist = cx.torch, cx.dist
    from dino_detector.losses import SetCriterion
    from dino_detector.matching import HungarianMatcher
    from dino_detector.optim import FusedAdam
    dev = cx.dev
    model = build_model(dev, "facebook/dinov2-large", lora_r=8).train()
    crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
    crit.strict = False
    # FusedAdam owns the flat fp32 gradient buffer (parallel.FlatGradSync): step() = all-reduce + clip + Adam
    opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    sync = opt.sync

    def images(n, seed):
        return torch.rand((n, 3, IMG, IMG), generator=torch.Generator(device=dev).manual_seed(seed), device=dev)

    def targets(n, seed):
        return [{k: v.to(dev) for k, v in t.items()} for t in make_targets(n, seed=seed)]

    parity = None
    if cx.world > 1:
        # ---- gradient parity, default switches (LayerNorm fold on), dropout 0, per-image matching ----
        p_drop, model.decoder.dropout_p = model.decoder.dropout_p, 0.0
        compat, crit.matcher.reference_compat = crit.matcher.reference_compat, False
        nb = parity_images
        xs = [images(nb, 1000 + r) for r in range(cx.world)]
        ts = [targets(nb, 2000 + r) for r in range(cx.world)]
        opt.zero_grad()
        sum(crit(model(xs[cx.rank]), ts[cx.rank]).values()).backward()
        sync.all_reduce(average=True)
        got = sync.flat.clone()
        crit.sync_num_boxes = False                  # single-process evaluation of the whole batch:
        sync.early = False                           # no collective from inside its backward either
        x_all, t_all = torch.cat(xs), [t for tt in ts for t in tt]

        def single():
            opt.zero_grad()
            sum(crit(model(x_all), t_all).values()).backward()
            return sync.flat.clone() / cx.world      # DDP averages; num_boxes is the global SUM on both sides

        want = single()
        again = single()                             # the same computation twice: what atomics alone move
        crit.sync_num_boxes = True
        sync.early = True
        names = {id(p): n for n, p in model.named_parameters()}

        def per_tensor(a_flat, b_flat):
            out, off = [], 0
            for p in sync.params:
                n = p.numel()
                a, b = a_flat[off:off + n], b_flat[off:off + n]
                out.append((float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)), names[id(p)], n))
                off += n
            return out

        rows, floor_rows = per_tensor(got, want), per_tensor(again, want)
        worst, worst_name, _ = max(rows)
        floor, floor_name, _ = max(floor_rows)
        # the two projections that produce sampling POSITIONS receive the part of the gradient that is a sum of
        # large cancelling terms (2- and 32-element bias tensors): their error is the fp32 atomics' order, see floor
        pos = ("reference_points_proj", "sampling_offsets")
        worst_other = max(r[0] for r in rows if not any(k in r[1] for k in pos))
        rels = [r[0] for r in rows]
        cos = float(torch.nn.functional.cosine_similarity(got, want, dim=0))
        stat = torch.tensor([worst, -cos, floor, worst_other], device=dev, dtype=torch.float64)
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)
        parity = {"worst_rel": float(stat[0]), "worst_tensor_rank0": worst_name,
                  "worst_rel_excluding_position_projections": float(stat[3]),
                  "nondeterminism_floor_worst_rel": float(stat[2]), "floor_tensor_rank0": floor_name,
                  "median_rel_rank0": statistics.median(rels), "cosine_min_over_ranks": -float(stat[1]),
                  "tensors": len(rels), "images_per_rank": nb, "ln_fold": os.environ.get("DOD_LN_FOLD", "1") != "0",
                  "note": "max |g_allreduced - g_single/N| / max |g_single/N| per parameter tensor; L/14 LoRA r=8 "
                          "deformable decoder, 518x518, dropout 0, per-image matching (reference_compat off so that "
                          "the concatenated batch pairs the same rows), num_boxes summed over ranks on both sides. "
                          "nondeterminism_floor = the same metric between two runs of the SAME single-process "
                          "computation (fp32 atomics of the deformable sampling backward land in another order)"}
        model.decoder.dropout_p, crit.matcher.reference_compat = p_drop, compat
        opt.zero_grad()
        del xs, ts, got, want, again, x_all
        torch.cuda.empty_cache()

    x = images(batch, 300 + cx.rank)
    tg = targets(batch, 400 + cx.rank)

    def step():
        opt.zero_grad()
        loss = sum(crit(model(x), tg).values())
        loss.backward()
        opt.step()                                   # all-reduce (NCCL) + clip + Adam
        return loss

    ms, _ = cx.timed(step, steps, warmup)
    gf = train_gflops("large", model.decoder.hidden_dim)
    rec = cx.rate_record(batch * cx.world, ms, steps, gf, warmup_steps=warmup, model="facebook/dinov2-large",
                         lora_r=8, decoder="deformable (reference default)", dropout=0.1,
                         trainable_params=sync.numel, grad_allreduce_bytes=sync.numel * 4,
                         collective="one NCCL all-reduce of the flat fp32 gradient per step + the 1-float num_boxes SUM, "
                                    "inside the timed step" + (" (DOD_EARLY_ALLREDUCE=1: projection + decoder part on a side "
                                                               "stream under the LoRA blocks' backward)"
                                                               if os.environ.get("DOD_EARLY_ALLREDUCE") == "1" else "")
                                    if cx.world > 1 else "none at N=1",
                         loss=float(step().detach()))
    if os.environ.get("DOD_BENCH_GRAPH", "1") != "0":
        # the same step replayed from ONE CUDA graph (runtime.GraphedTrainStep: forward, GPU matcher, fused criterion,
        # backward, the NCCL all-reduces, clip + Adam; targets re-staged from the host every step): what the ~700
        # launch gaps of the eager step cost
        try:
            from dino_detector.runtime import GraphedTrainStep
            tg_host = make_targets(batch, seed=400 + cx.rank)
            gstep = GraphedTrainStep(model, crit, opt, x, max_targets=max(t["labels"].numel() for t in tg_host) + 1)
            gms, _ = cx.timed(lambda: gstep(x, tg_host), steps, warmup)
            rec["graph"] = {"ms_per_step": gms / steps, "value": batch * cx.world * steps / (gms * 1e-3), "unit": UNIT,
                            "loss": float(sum(gstep(x, tg_host).values())),
                            "what": "the same train step as ONE CUDA graph replay per step (NCCL all-reduces captured)"}
            del gstep
        except Exception as e:                                   # never take the bench line down
            rec["graph"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    del model, opt, x
    torch.cuda.empty_cache()
    return rec, parity


def measure_torch_eager(cx, steps, warmup):
    """The same-box bar: the UNMODIFIED reference module (baseline/_ref) on this GPU under torch.autocast(bf16),
    standard decoder (the deformable reference is host-sync-bound: ~2e5 .item() calls per layer at batch 64)."""
    torch = cx.torch
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref is missing"}
    torch.manual_seed(0)
    model = ref_loader.build_detector(dino_model_name=MODEL_NAME, use_deformable=False).to(cx.dev).eval()
    x = torch.rand((BATCH, 3, IMG, IMG), generator=torch.Generator().manual_seed(0)).to(cx.dev)

    def step():
        model(x)

    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ms, _ = cx.timed(step, steps, warmup)
    rec = {"value": BATCH * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "what": "unmodified reference DINOv2ObjectDetector (baseline/_ref; HF Dinov2Model + nn.TransformerDecoder, "
                   "use_deformable=False), PyTorch eager, torch.autocast(bf16), same 64 x 518x518 batch, this GPU",
           "torch": torch.__version__}
    del model, x
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    cx = Ctx()
    torch = cx.torch
    sampler = None
    if cx.rank == 0:
        sampler = ClockSampler(cx.local)
        sampler.start()
    if args.gpus != cx.world and cx.rank == 0 and cx.world > 1:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={cx.world}; using {cx.world}", file=sys.stderr)

    head = measure_headline(cx, args, sampler)
    records = {}
    if not args.no_extra:
        k = max(2, min(args.steps, 8))
        records["infer_l14"] = measure_inference(cx, "facebook/dinov2-large", 64, 64, k, 3, "large")
        per_gpu = max(32, 512 // cx.world)
        records["infer_c5"] = measure_inference(cx, "facebook/dinov2-giant", per_gpu, 64, 2 if per_gpu > 128 else 3,
                                                1 if per_gpu > 128 else 2, "giant")
        records["infer_c5"]["global_batch"] = per_gpu * cx.world
        rec, parity = measure_train(cx, k, 3)
        records["train_c4"] = rec
        if parity is not None:
            records["ddp_parity"] = parity
        if cx.world == 1:
            try:
                records["torch_eager_b200"] = measure_torch_eager(cx, max(2, min(args.steps, 5)), 3)
            except Exception as e:                               # the bar must never take the bench line down
                records["torch_eager_b200"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if sampler is not None:
        sampler.stop()

    if cx.rank != 0:
        if cx.world > 1:
            cx.dist.destroy_process_group()
        return

    peaks = cx.peaks
    world, steps = cx.world, args.steps
    ms_dev = head["ms_dev"]
    n_img = BATCH * world * steps
    value = n_img / (ms_dev * 1e-3)
    prof = head["prof"]
    gl, gf, gt = prof.get("gemm", (0, 0.0, 1e-9))
    fl, ff, ft = prof.get("fmha", (0, 0.0, 1e-9))
    ll, lb, lt = prof.get("layernorm", (0, 0.0, 1e-9))
    gemm_tflops = gf / (gt * 1e-3) / 1e12
    fmha_tflops = ff / (ft * 1e-3) / 1e12
    fwd = forward_gflops("base")
    step_tflops = fwd["total"] * BATCH / (head["ms_local"] / steps)            # GF / ms = TFLOP/s
    roof = {"kernel": "dod::gemm2_kernel / gemm_kernel (tcgen05 cta_group::2, TMEM, TMA)", "bound": "tensor",
            "achieved": gemm_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": gemm_tflops / peaks["tflops"],
            "frac_of_burst": gemm_tflops / peaks["tflops_burst"], "peak_burst": peaks["tflops_burst"],
            "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step); burst beside it",
            "traffic": _ncu_traffic(), "launches_per_step": gl / steps,
            "algorithmic_gflops_per_launch": gf / max(gl, 1) / 1e9, "avg_launch_us": 1e3 * gt / max(gl, 1),
            "share_of_step": gt / head["ms_local"]}
    extra = {
        "fmha": {"kernel": "dod::fmha_kernel", "bound": "tensor", "achieved": fmha_tflops,
                 "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": fmha_tflops / peaks["tflops"],
                 "frac_of_burst": fmha_tflops / peaks["tflops_burst"], "launches_per_step": fl / steps,
                 "avg_launch_us": 1e3 * ft / max(fl, 1), "share_of_step": ft / head["ms_local"]},
        "layernorm": {"kernel": "dod::layernorm_kernel", "bound": "hbm", "achieved": lb / (lt * 1e-3) / 1e9,
                      "peak": peaks["hbm"], "unit": "GB/s", "frac": lb / (lt * 1e-3) / 1e9 / peaks["hbm"],
                      "launches_per_step": ll / steps, "share_of_step": lt / head["ms_local"]},
        "whole_step": {"bound": "tensor", "achieved": step_tflops, "unit": "TFLOP/s",
                       "gflops_per_image": round(fwd["total"], 2), "frac": step_tflops / peaks["tflops"],
                       "frac_of_burst": step_tflops / peaks["tflops_burst"]},
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        try:
            rate, times = reference_forward_rate(REF_SAMPLE_IMAGES, 12, 1, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{REF_SAMPLE_IMAGES} of {BATCH} images per forward, {len(times)} timed forwards of the unmodified "
                             f"reference detector (baseline/_ref, default ctor, fp32, eval, no_grad), median "
                             f"{statistics.median(times):.2f} s, total {sum(times):.1f} s"}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"unavailable: {type(e).__name__}: {e}"[:300]}
    e2e, e2u = head["e2e"], head["e2e_uint8"]
    records["e2e_uint8"] = {"value": n_img / (e2u["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e2u["ms"] / steps,
                            "h2d_bytes_per_step": e2u["h2d"], "d2h_bytes_per_step": e2u["d2h"],
                            "note": "same workload, images as uint8 [B, H, W, 3] in pinned host memory (ToTensor's "
                                    "/255 is fused into the im2col kernel)"}
    if "torch_eager_b200" in records and records["torch_eager_b200"].get("value"):
        records["torch_eager_b200"]["ours_over_eager"] = value / records["torch_eager_b200"]["value"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(world),
        "roofline": roof, "roofline_other": extra, "cpu_baseline": cpu,
        "e2e": {"value": n_img / (e2e["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["ms"] / steps,
                "note": "pinned host fp32 images, double-buffered H2D on a copy stream, outputs copied back every step"},
        "gpu_launches": int(head["launches"]), "clocks": head["clocks"], "records": records,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        cx.dist.destroy_process_group()


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (skip the `records` of the other configs)")
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
