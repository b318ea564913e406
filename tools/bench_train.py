"""C4: L/14, LoRA r=8, deformable decoder, full train step (forward + matcher + fused
criterion + backward + gradient all-reduce + clip + Adam) at 518x518, data parallel over N GPUs.

    python tools/bench_train.py [--batch 32] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29511 tools/bench_train.py --batch 32

One process per GPU, per-GPU batch fixed (weak scaling); the only exchange step is ONE all-reduce of
the flat trainable-gradient buffer (parallel.FlatGradSync) + the 1-float num_boxes SUM inside the
criterion (losses.py:228-230).  Timed on the device (CUDA events), max over ranks; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dinov2-od_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="facebook/dinov2-large")
    ap.add_argument("--lora-r", type=int, default=8)
    ap.add_argument("--hw", type=int, default=518, help="image side")
    ap.add_argument("--graph", action="store_true", help="replay the step from one CUDA graph (runtime.GraphedTrainStep)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import _inputs as synth
    from dino_detector.models import DINOv2ObjectDetector
    from dino_detector.losses import SetCriterion
    from dino_detector.matching import HungarianMatcher
    from dino_detector.optim import FusedAdam

    torch.manual_seed(0)
    model = DINOv2ObjectDetector(dino_model_name=a.model, lora_r=a.lora_r).cuda().train()
    crit = SetCriterion(HungarianMatcher(), 91, {"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0})
    # FusedAdam owns the flat fp32 gradient buffer (parallel.FlatGradSync): step() = all-reduce + clip + Adam
    opt = FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0)
    sync = opt.sync
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(a.batch, 3, a.hw, a.hw, generator=g).cuda()
    targets = [{k: v.cuda() for k, v in t.items()}
               for t in synth.make_targets(a.batch, max_gt=20, seed=3 + rank, min_gt=1)]

    def step():
        opt.zero_grad()
        loss = sum(crit(model(x), targets).values())
        loss.backward()
        opt.step()
        return loss

    if a.graph:
        from dino_detector.runtime import GraphedTrainStep
        graphed = GraphedTrainStep(model, crit, opt, x, max_targets=20)
        host_targets = [{k: v.cpu() for k, v in t.items()} for t in targets]

        def step():  # noqa: F811  (refills the static image / target buffers, then one graph launch)
            ld = graphed(x, host_targets)
            return ld["loss_ce"] + ld["loss_bbox"] + ld["loss_giou"]

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": f"c4 {a.model} LoRA r={a.lora_r} + deformable decoder train step, bf16, {a.hw}x{a.hw}",
                          "n_gpus": world, "batch_per_gpu": a.batch, "cuda_graph": bool(a.graph), "ms_per_step": ms.item(),
                          "images_per_s": world * a.batch / (ms.item() * 1e-3), "scaling": "weak",
                          "loss": float(loss), "trainable_params": sync.numel,
                          "grad_allreduce_bytes": sync.numel * 4}))
    if world > 1:
        if a.graph:
            # tearing the communicator down while captured graphs still hold NCCL kernels hangs in
            # destroy_process_group (torch 2.11 / NCCL 2.28): leave together and skip the teardown
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
