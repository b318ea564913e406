"""CPU, world_size 2 over gloo: the data-parallel host logic (batch sharding, flat gradient
all-reduce == single-process gradient of the concatenated batch, num_boxes semantics)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (sets sys.path)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dino_detector.parallel import FlatGradSync, all_reduce_num_boxes, shard_range
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    model[0].bias.requires_grad_(False)                      # frozen tensors are skipped
    x = torch.randn(10, 16)
    y = torch.randn(10, 3)
    # single-process reference on the whole batch (mean over the batch)
    ref = torch.autograd.grad(((model(x) - y) ** 2).mean(), [p for p in model.parameters() if p.requires_grad])
    sync = FlatGradSync(model.parameters())
    s, e = shard_range(10, rank, world)
    sync.zero()
    # per-rank mean over its shard, weighted so that the rank average equals the global mean
    loss = ((model(x[s:e]) - y[s:e]) ** 2).sum() / (10 * 3) * world
    loss.backward()
    sync.all_reduce(average=True)
    ok = all(torch.allclose(p.grad, g, atol=1e-6) for p, g in zip(sync.params, ref))
    nb = all_reduce_num_boxes(torch.tensor([float(rank + 1)]))
    results[rank] = (ok, float(nb.item()), (s, e), sync.numel)
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert results[0][0] and results[1][0]
    assert results[0][1] == 3.0 and results[1][1] == 3.0      # SUM, not averaged (losses.py:228-229)
    assert results[0][2] == (0, 5) and results[1][2] == (5, 10)
    assert results[0][3] == 16 * 8 + 8 * 3 + 3


def test_shard_range_covers_everything():
    from dino_detector.parallel import shard_range
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_zero_grad_set_to_none_is_detected():
    from dino_detector.parallel import FlatGradSync
    lin = torch.nn.Linear(4, 4)
    sync = FlatGradSync(lin.parameters())
    lin.zero_grad(set_to_none=True)
    sync.attach()
    assert lin.weight.grad.data_ptr() == sync.flat.data_ptr()
