"""2-GPU (NCCL) data-parallel training parity, SURVEY.md 8d C4: gradients after the DDP all-reduce
equal the single-process gradients on the concatenated batch.  Needs >= 2 CUDA devices
(`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss(out):
    return torch.nn.functional.softplus(out["pred_logits"]).sum() + ((out["pred_boxes"] - 0.3) ** 2).sum()


def _worker(rank, world, port, results, use_ddp):
    import sys
    micro = 1
    if use_ddp in ("flat-early", "flat-early-accum"):
        os.environ["DOD_EARLY_ALLREDUCE"] = "1"      # projection + decoder part all-reduced from inside the backward
        micro = 2 if use_ddp == "flat-early-accum" else 1     # two backward passes before the optimizer's all-reduce
        use_ddp = False
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import build_product_model, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, sd, kw = build_product_model("c1_small_std", device=f"cuda:{rank}", dropout=0.0)
    model.train()
    x = synth.make_images(world * micro, 224, 224, seed=21)
    if use_ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        ddp = DDP(model, device_ids=[rank], find_unused_parameters=True)      # reference train.py:677
        _loss(ddp(x[rank:rank + 1].cuda())).backward()
    else:
        from dino_detector.parallel import FlatGradSync
        sync = FlatGradSync(model.parameters())
        sync.zero()
        for k in range(micro):
            i = k * world + rank
            _loss(model(x[i:i + 1].cuda())).backward()
        if os.environ.get("DOD_EARLY_ALLREDUCE") == "1":
            assert sync._early is not None and sync._early[1] > 0, "the early all-reduce did not start in the backward"
        sync.all_reduce(average=True)
    torch.cuda.synchronize()
    if rank == 0:
        avg = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        ref_model, _, _ = build_product_model("c1_small_std", device="cuda:0", dropout=0.0)
        ref_model.train()
        _loss(ref_model(x.cuda())).backward()
        worst = 0.0
        for n, p in ref_model.named_parameters():
            if p.grad is None:
                continue
            want = p.grad / world                                  # DDP averages, the sum-loss does not
            err = ((avg[n] - want).abs().max() / want.abs().max().clamp_min(1e-9)).item()
            worst = max(worst, err)
        results["worst"] = worst
        results["n"] = len(avg)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_ddp", [True, False, "flat-early", "flat-early-accum"])
def test_two_gpu_gradient_allreduce_matches_single_process(use_ddp):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    results = mgr.dict()
    # default switches (LayerNorm fold on): whether a LayerNorm is folded no longer depends on the number of token
    # rows, so the one-image shards and the concatenated two-image batch run the same arithmetic per image
    mp.spawn(_worker, args=(2, _free_port(), results, use_ddp), nprocs=2, join=True)
    assert results["n"] > 50
    assert results["worst"] < 2e-2, results["worst"]
